"""Host logic of the enhanced-recipe training driver (DM/train_dm.py:65-71,121-226): the curriculum state machine,
the mt-ratio schedule and the flag surface - no GPU, no compute calls."""
import pytest

from open_l2o_b200.train_dm import NUM_STEPS, Curriculum, build_parser


def test_curriculum_tables_follow_the_reference():
    c = Curriculum(unroll_length=20)
    assert c.num_steps == [100, 200, 500, 1000, 1500, 2000, 2500, 3000] == NUM_STEPS      # DM/train_dm.py:66
    assert c.num_unrolls == [5, 10, 25, 50, 75, 100, 125, 150]                            # :67
    assert c.num_unrolls_eval == c.num_unrolls[1:]                                        # :68
    assert c.train_unrolls() == 5 and c.eval_unrolls() == 10
    assert c.mt_ratio([0.3, 0.2, 0.1]) == 0.3


def test_curriculum_save_advance_stop_sequence():
    """DM/train_dm.py:177-222: a new best saves; after >= min_num_eval evaluations with an improvement the driver
    restores the best model and advances; >= min_num_eval without any improvement stops."""
    c = Curriculum(unroll_length=20, min_num_eval=3)
    assert c.observe(5.0) == ("save", 0)
    assert c.observe(4.0) == ("save", 0)
    assert c.observe(4.5) == ("advance", 0, 1)          # 3rd evaluation, improved earlier
    assert (c.idx, c.num_eval, c.improved) == (1, 0, False)
    c.rebase(3.9)                                        # the driver's fresh evaluation at the new curriculum
    assert c.train_unrolls() == 10 and c.eval_unrolls() == 25
    assert c.observe(4.2) == ("continue", 1)
    assert c.observe(4.1) == ("continue", 1)
    assert c.observe(4.0) == ("stop", 1)                # 3 evaluations, none better than 3.9
    assert c.mt_ratio([0.3]) == 0.3                      # idx past the list -> last ratio (:124-127)


def test_curriculum_wraps_to_minus_one_after_the_last_stage():
    c = Curriculum(unroll_length=20, min_num_eval=1)
    for stage in range(len(c.num_unrolls)):
        assert c.observe(10.0 - stage) [0] == "save"
        act = c.observe(100.0)
        assert act[0] == "advance"
        c.rebase(50.0)
    assert c.idx == -1                                   # DM/train_dm.py:203-204
    assert c.train_unrolls() == c.num_unrolls[-1] and c.eval_unrolls() == c.num_unrolls_eval[-1]


def test_flag_surface_matches_the_reference_drivers():
    """Flag names and defaults of DM/train_dm.py:33-60 / DM/train_rnnprop.py:33-60."""
    a = build_parser().parse_args([])
    assert (a.num_epochs, a.evaluation_period, a.evaluation_epochs, a.num_steps, a.unroll_length) == (10000, 100, 20, 100, 20)
    assert a.learning_rate == 0.001 and a.second_derivatives is False
    assert (a.if_scale, a.rd_scale_bound, a.if_cl, a.min_num_eval) == (False, 3.0, False, 3)
    assert (a.if_mt, a.num_mt, a.optimizers, a.mt_ratio, a.mt_ratios, a.k) == (False, 1, "adam", 0.3, None, 1)
    assert (a.beta1, a.beta2) == (0.95, 0.95)
    with pytest.raises(SystemExit):
        build_parser().parse_args(["--net", "sgd"])
