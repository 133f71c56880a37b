"""Regenerates tests/golden/bias_gru_hand.json: single-cell BiasGRUCell known answers computed with scalar `math` only
(no torch, no oracle import) from the reference's code as written (SC/optimizer/rnn_cells.py:46-68 with
SC/optimizer/utils.py:36-90 `affine` = concat(inputs) @ Matrix + Bias):

    bias -> (r_bias | u_bias | c_bias)                       three equal column blocks
    [r_lin | u_lin] = [x | h] Wg + bg ;  r = sigmoid(r_lin + r_bias), u = sigmoid(u_lin + u_bias)
    c = tanh([x | r*h] Wc + bc + c_bias) ;  h' = u*h + (1-u)*c

The INPUTS below are fixed literals (not read back from the file this script writes), chosen to pin what a wrong
restatement would get wrong: gate order (r before u), which gate multiplies the old state, the reset gate acting on h
BEFORE the candidate affine, the injected bias split order, and a per-row (broadcast) injected bias.

    python tests/golden/make_bias_gru_hand.py          # rewrites the file
"""
import json
import math
import os

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bias_gru_hand.json")


def sigmoid(v):
    return 1.0 / (1.0 + math.exp(-v))


def cell(x, h, wg, bg, wc, bc, bias):
    n, hs = len(x), len(h[0])
    out = []
    for r_ in range(n):
        b = bias[r_ if len(bias) > 1 else 0]
        rb, ub, cb = b[0:hs], b[hs:2 * hs], b[2 * hs:3 * hs]
        row = list(x[r_]) + list(h[r_])
        proj = [sum(row[k] * wg[k][col] for k in range(len(row))) + bg[col] for col in range(2 * hs)]
        r = [sigmoid(proj[u] + rb[u]) for u in range(hs)]
        u = [sigmoid(proj[hs + k] + ub[k]) for k in range(hs)]
        row2 = list(x[r_]) + [r[k] * h[r_][k] for k in range(hs)]
        c = [math.tanh(sum(row2[k] * wc[k][col] for k in range(len(row2))) + bc[col] + cb[col]) for col in range(hs)]
        out.append([u[k] * h[r_][k] + (1.0 - u[k]) * c[k] for k in range(hs)])
    return out


def cases():
    out = []
    # 1. all-zero weights: r = u = 1/2, c = 0 -> h' = h/2 (closed form)
    out.append(dict(name="zero_weights_halve_state", x=[[0.7, -1.2]], h=[[0.4, -0.8]],
                    wg=[[0.0] * 4 for _ in range(4)], bg=[0.0] * 4, wc=[[0.0] * 2 for _ in range(4)], bc=[0.0, 0.0],
                    bias=[[0.0] * 6]))
    # 2. huge update-gate bias: u -> 1, h' = h whatever the candidate is; huge NEGATIVE: h' = c
    out.append(dict(name="update_gate_keeps_state", x=[[0.3, 0.9]], h=[[0.25, -0.5]],
                    wg=[[0.1, -0.2, 0.3, 0.4], [0.0, 0.1, -0.1, 0.2], [0.5, 0.5, 0.0, 0.0], [-0.3, 0.2, 0.1, 0.1]],
                    bg=[0.0, 0.0, 40.0, 40.0], wc=[[0.2, -0.1], [0.4, 0.3], [0.6, -0.6], [0.1, 0.9]], bc=[0.05, -0.05],
                    bias=[[0.0] * 6]))
    out.append(dict(name="update_gate_takes_candidate", x=[[0.3, 0.9]], h=[[0.25, -0.5]],
                    wg=[[0.1, -0.2, 0.3, 0.4], [0.0, 0.1, -0.1, 0.2], [0.5, 0.5, 0.0, 0.0], [-0.3, 0.2, 0.1, 0.1]],
                    bg=[0.0, 0.0, -40.0, -40.0], wc=[[0.2, -0.1], [0.4, 0.3], [0.6, -0.6], [0.1, 0.9]], bc=[0.05, -0.05],
                    bias=[[0.0] * 6]))
    # 3. reset gate shut (r -> 0) through the INJECTED r_bias only: the candidate must ignore h entirely
    out.append(dict(name="injected_reset_bias_blocks_state", x=[[1.0, -2.0]], h=[[5.0, -7.0]],
                    wg=[[0.0] * 4 for _ in range(4)], bg=[0.0] * 4,
                    wc=[[0.3, 0.0], [0.0, 0.2], [1.0, 1.0], [1.0, -1.0]], bc=[0.0, 0.0],
                    bias=[[-50.0, -50.0, 0.0, 0.0, 0.1, -0.1]]))
    # 4. general case, 3 rows, per-row injected bias, asymmetric everything
    out.append(dict(name="general_three_rows_per_row_bias",
                    x=[[0.5, -1.0, 0.25], [-0.75, 0.1, 2.0], [0.0, 0.0, 0.0]],
                    h=[[0.1, -0.2], [0.9, 0.3], [-0.6, 0.6]],
                    wg=[[0.11, -0.21, 0.31, -0.41], [0.52, 0.62, -0.72, 0.82], [-0.13, 0.23, 0.33, -0.43],
                        [0.74, -0.64, 0.54, 0.44], [-0.35, 0.25, -0.15, 0.05]],
                    bg=[2.2, 2.2, 2.2, 2.2],     # gate_bias_init of the reference's drivers (SC/metarun.py)
                    wc=[[0.6, -0.5], [0.4, 0.3], [-0.2, 0.1], [0.9, -0.8], [0.7, 0.65]], bc=[0.01, -0.02],
                    bias=[[0.3, -0.3, 0.2, -0.2, 0.1, -0.1], [-1.0, 1.0, 0.5, -0.5, 0.0, 0.25], [0.0] * 6]))
    # 5. one shared (1-row) injected bias broadcast over two rows (the per-tensor / global GRUs' use)
    out.append(dict(name="broadcast_bias_two_rows", x=[[0.2], [-0.4]], h=[[0.3, 0.1, -0.2], [0.0, 0.5, 0.5]],
                    wg=[[0.1, 0.2, 0.3, -0.1, -0.2, -0.3], [0.4, 0.0, -0.4, 0.2, 0.0, -0.2],
                        [0.0, 0.3, 0.0, -0.3, 0.0, 0.3], [0.25, -0.25, 0.5, -0.5, 0.75, -0.75]],
                    bg=[0.1, 0.2, 0.3, 0.4, 0.5, 0.6],
                    wc=[[0.5, -0.5, 0.25], [0.1, 0.2, 0.3], [-0.3, -0.2, -0.1], [0.7, 0.0, -0.7]], bc=[0.0, 0.1, -0.1],
                    bias=[[0.05, -0.05, 0.15, -0.15, 0.25, -0.25, 0.35, -0.35, 0.45]]))
    for c in out:
        c["h_next"] = cell(c["x"], c["h"], c["wg"], c["bg"], c["wc"], c["bc"], c["bias"])
    return out


if __name__ == "__main__":
    json.dump(cases(), open(PATH, "w"), indent=1)
    print("wrote", PATH, [c["name"] for c in cases()])
