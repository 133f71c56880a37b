"""The enhanced-recipe training driver (DM/train_dm.py and its twin DM/train_rnnprop.py): curriculum over the number
of optimization steps per epoch, imitation-task ("mt") mixing with ``mt_ratio``, random scaling, periodic evaluation
with best-model save / restore of ``<net>.l2l-<idx>``.  SURVEY.md 8(f) row 2 - this is the caller that makes the fused
unroll regime the common case; the schedule itself is host logic (``Curriculum``, testable without a GPU).

    python -m open_l2o_b200.train_dm --problem quadratic --if_cl --if_mt --num_mt 1 --optimizers adam --save_path /tmp/o
"""
from __future__ import annotations

import argparse
import os
import random
from timeit import default_timer as timer

NUM_STEPS = [100, 200, 500, 1000, 1500, 2000, 2500, 3000]   # DM/train_dm.py:66


class Curriculum(object):
    """The schedule state machine of DM/train_dm.py:65-71,177-222.  ``observe(eval_cost)`` is called after every
    evaluation and returns the action the driver must take:

      ("save", idx)                 new best inside curriculum ``idx``: save ``.l2l-idx`` and ``.l2l-0``
      ("advance", old_idx, new_idx) >= min_num_eval evaluations and an improvement was seen: restore ``.l2l-old_idx``,
                                    move to ``new_idx``; the driver re-evaluates and reports it with ``rebase(cost)``
      ("stop", idx)                 >= min_num_eval evaluations without any improvement
      ("continue", idx)             keep training
    """

    def __init__(self, unroll_length, min_num_eval=3, num_steps=None):
        self.num_steps = list(num_steps or NUM_STEPS)
        self.num_unrolls = [int(ns / unroll_length) for ns in self.num_steps]
        self.num_unrolls_eval = self.num_unrolls[1:]
        self.min_num_eval = min_num_eval
        self.idx = 0
        self.best = float("inf")
        self.num_eval = 0
        self.improved = False

    def train_unrolls(self):
        return self.num_unrolls[self.idx]

    def eval_unrolls(self):
        return self.num_unrolls_eval[self.idx]

    def mt_ratio(self, ratios):
        return ratios[-1] if self.idx >= len(ratios) else ratios[self.idx]

    def observe(self, eval_cost):
        self.num_eval += 1
        if eval_cost < self.best:
            self.best = eval_cost
            self.improved = True
            return ("save", self.idx)
        if self.num_eval >= self.min_num_eval and self.improved:
            old = self.idx
            self.num_eval = 0
            self.improved = False
            self.idx += 1
            if self.idx >= len(self.num_unrolls):
                self.idx = -1                      # DM/train_dm.py:203-204
            return ("advance", old, self.idx)
        if self.num_eval >= self.min_num_eval and not self.improved:
            return ("stop", self.idx)
        return ("continue", self.idx)

    def rebase(self, eval_cost):
        self.best = eval_cost


def build_parser():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--save_path", default=None)
    ap.add_argument("--num_epochs", type=int, default=10000)
    ap.add_argument("--evaluation_period", type=int, default=100)
    ap.add_argument("--evaluation_epochs", type=int, default=20)
    ap.add_argument("--num_steps", type=int, default=100)
    ap.add_argument("--unroll_length", type=int, default=20)
    ap.add_argument("--learning_rate", type=float, default=0.001)
    ap.add_argument("--second_derivatives", action="store_true")
    ap.add_argument("--problem", default="quadratic")
    ap.add_argument("--net", default="dm", choices=["dm", "rnnprop"], help="dm = train_dm.py, rnnprop = train_rnnprop.py")
    ap.add_argument("--beta1", type=float, default=0.95)
    ap.add_argument("--beta2", type=float, default=0.95)
    ap.add_argument("--if_scale", action="store_true")
    ap.add_argument("--rd_scale_bound", type=float, default=3.0)
    ap.add_argument("--if_cl", action="store_true")
    ap.add_argument("--min_num_eval", type=int, default=3)
    ap.add_argument("--if_mt", action="store_true")
    ap.add_argument("--num_mt", type=int, default=1)
    ap.add_argument("--optimizers", default="adam")
    ap.add_argument("--mt_ratio", type=float, default=0.3)
    ap.add_argument("--mt_ratios", default=None,
                    help='per-curriculum mt ratios; default "0.0 0.1 0.3 0.3 0.3 0.3 0.3 0.3" (train_dm.py:59) or '
                         '"0.3 0.3 0.3" with --net rnnprop (train_rnnprop.py)')
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    return ap


def train(FLAGS, problem=None, net_config=None, net_assignments=None, log=print):
    """DM/train_dm.py:63-226 (``--net rnnprop``: DM/train_rnnprop.py).  Returns a dict with the per-evaluation history."""
    from . import meta_dm_train, meta_rnnprop_train, util
    from .data_generator import data_loader
    from .meta import Session

    rnnprop = FLAGS.net == "rnnprop"
    cl = Curriculum(FLAGS.unroll_length, FLAGS.min_num_eval) if FLAGS.if_cl else None
    num_unrolls = FLAGS.num_steps // FLAGS.unroll_length
    if FLAGS.save_path is not None and not os.path.exists(FLAGS.save_path):
        os.mkdir(FLAGS.save_path)
    if problem is None:
        problem, net_config, net_assignments = util.get_config(FLAGS.problem, net_name="RNNprop" if rnnprop else None)
    rng = random.Random(FLAGS.seed)

    kw = dict(learning_rate=FLAGS.learning_rate, net_assignments=net_assignments,
              second_derivatives=FLAGS.second_derivatives)
    seq_step = None
    if rnnprop:
        optimizer = meta_rnnprop_train.MetaOptimizer(FLAGS.num_mt, FLAGS.beta1, FLAGS.beta2, _seed=FLAGS.seed, **net_config)
        (minimize, scale, var_x, constants, subsets, seq_step,
         loss_mt, steps_mt, update_mt, reset_mt, mt_labels, mt_inputs) = optimizer.meta_minimize(
            problem, FLAGS.unroll_length, **kw)
    else:
        optimizer = meta_dm_train.MetaOptimizer(FLAGS.num_mt, _seed=FLAGS.seed, **net_config)
        (minimize, scale, var_x, constants, subsets,
         loss_mt, steps_mt, update_mt, reset_mt, mt_labels, mt_inputs) = optimizer.meta_minimize(
            problem, FLAGS.unroll_length, **kw)
    optimizer.restorer()
    step, update, reset, cost_op, _ = minimize
    step_kw = dict(step=seq_step, unroll_len=FLAGS.unroll_length) if rnnprop else {}
    data_mt = None
    if FLAGS.if_mt:
        data_mt = data_loader(problem, var_x, constants, subsets, scale, FLAGS.optimizers, FLAGS.unroll_length)
    mt_ratios = [float(r) for r in (FLAGS.mt_ratios or ("0.3 0.3 0.3" if rnnprop else
                                                         "0.0 0.1 0.3 0.3 0.3 0.3 0.3 0.3")).split()]
    assign_func = optimizer.assign_func

    history = []
    with Session() as sess:
        for rst in [reset] + list(reset_mt):
            sess.run(rst)
        start_time = timer()
        best_evaluation = float("inf")
        mti = -1

        def evaluate(n_unrolls):
            tot = 0.0
            for _ in range(FLAGS.evaluation_epochs):
                _, cost = util.run_epoch(sess, cost_op, [update], reset, n_unrolls, **step_kw)
                tot += cost
            return tot

        for e in range(FLAGS.num_epochs):
            task_i = -1
            if FLAGS.if_mt:                                             # DM/train_dm.py:121-135
                mt_ratio = cl.mt_ratio(mt_ratios) if cl is not None else FLAGS.mt_ratio
                if rng.random() < mt_ratio:
                    mti = (mti + 1) % FLAGS.num_mt
                    task_i = mti
            num_unrolls_cur = cl.train_unrolls() if cl is not None else num_unrolls
            if task_i == -1:
                _, cost = util.run_epoch(sess, cost_op, [update, step], reset, num_unrolls_cur, scale=scale,
                                         rd_scale=FLAGS.if_scale, rd_scale_bound=FLAGS.rd_scale_bound,
                                         assign_func=assign_func, var_x=var_x, **step_kw)
            else:
                data_e = data_mt.get_data(task_i, sess, num_unrolls_cur, assign_func, FLAGS.rd_scale_bound,
                                          if_scale=FLAGS.if_scale, mt_k=FLAGS.k)
                _, cost = util.run_epoch(sess, loss_mt[task_i], [update_mt[task_i], steps_mt[task_i]], reset_mt[task_i],
                                         num_unrolls_cur, task_i=task_i, data=data_e, label_pl=mt_labels[task_i],
                                         input_pl=mt_inputs[task_i], **step_kw)
            log("training_loss={}".format(cost))

            if (e + 1) % FLAGS.evaluation_period != 0:
                continue
            eval_cost = evaluate(cl.eval_unrolls() if cl is not None else num_unrolls)
            num_steps_cur = cl.num_steps[cl.idx] if cl is not None else FLAGS.num_steps
            log("epoch={}, num_steps={}, eval_loss={}".format(e, num_steps_cur, eval_cost / FLAGS.evaluation_epochs))
            history.append(dict(epoch=e, num_steps=num_steps_cur, eval_loss=eval_cost / FLAGS.evaluation_epochs,
                                task=task_i))
            if cl is None:                                              # DM/train_dm.py:167-173
                if eval_cost < best_evaluation:
                    best_evaluation = eval_cost
                    if FLAGS.save_path is not None:
                        optimizer.save(sess, FLAGS.save_path, e + 1)
                        optimizer.save(sess, FLAGS.save_path, 0)
                        log("Saving optimizer of epoch {}...".format(e + 1))
                continue
            action = cl.observe(eval_cost)                              # DM/train_dm.py:175-222
            history[-1]["action"] = action[0]
            if action[0] == "save":
                if FLAGS.save_path is not None:
                    optimizer.save(sess, FLAGS.save_path, action[1])
                    optimizer.save(sess, FLAGS.save_path, 0)
            elif action[0] == "advance":
                if FLAGS.save_path is not None:
                    optimizer.restore(sess, FLAGS.save_path, action[1])
                eval_cost = evaluate(cl.eval_unrolls())
                cl.rebase(eval_cost)
                log("epoch={}, num_steps={}, eval loss={}".format(e, cl.num_steps[cl.idx],
                                                                 eval_cost / FLAGS.evaluation_epochs))
            elif action[0] == "stop":
                log("no improve during curriculum {} --> stop".format(action[1]))
                break
        log("total time = {}s...".format(timer() - start_time))
    return dict(history=history, optimizer=optimizer)


def main(argv=None):
    train(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
