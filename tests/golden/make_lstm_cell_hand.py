"""Regenerates tests/golden/lstm_cell_hand.json: single-cell LSTM known answers computed with scalar `math` only
(no torch, no oracle import) from the Sonnet-1.11 recipe the reference relies on (DM/networks.py:197 -> snt.LSTM):
z = [x | h] W + b, columns split i | j | f | o, c' = sigmoid(f + 1) c + sigmoid(i) tanh(j), h' = tanh(c') sigmoid(o).
The inputs are read from the committed file, only h_next / c_next are recomputed.

    python tests/golden/make_lstm_cell_hand.py          # rewrites the file in place
"""
import json
import math
import os

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lstm_cell_hand.json")


def sigmoid(v):
    return 1.0 / (1.0 + math.exp(-v))


def cell(x, h, c, w, b):
    n, hs = len(x), len(h[0])
    hn, cn = [], []
    for r in range(n):
        row = list(x[r]) + list(h[r])
        z = [sum(row[k] * w[k][col] for k in range(len(row))) + b[col] for col in range(4 * hs)]
        i, j, f, o = z[0:hs], z[hs:2 * hs], z[2 * hs:3 * hs], z[3 * hs:4 * hs]
        c_row = [sigmoid(f[u] + 1.0) * c[r][u] + sigmoid(i[u]) * math.tanh(j[u]) for u in range(hs)]
        hn.append([math.tanh(c_row[u]) * sigmoid(o[u]) for u in range(hs)])
        cn.append(c_row)
    return hn, cn


def regenerate():
    cases = json.load(open(PATH))
    for cse in cases:
        cse["h_next"], cse["c_next"] = cell(cse["x"], cse["h"], cse["c"], cse["w"], cse["b"])
    return cases


if __name__ == "__main__":
    out = regenerate()
    json.dump(out, open(PATH, "w"), indent=1)
    print("wrote", PATH, [c["name"] for c in out])
