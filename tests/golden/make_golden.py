"""Generates tests/golden/sage_known_answer.json — an exact-arithmetic known answer for SAGEConv.

Independent of torch and of oracle/: python Fractions over a hand-written 6-node / 9-edge graph with
one zero-in-degree node (5), one duplicated edge (1->0 twice), one self loop (2->2) and unsorted
edge order.  Semantics per SURVEY §8 A5: out_i = W_l * mean_{j->i} x_j + b_l + W_r * x_i, mean over
max(count,1), duplicates counted, no self loops added.  Also the exact gradients of
loss = sum_i sum_o c[i][o] * out[i][o] w.r.t. W_l, W_r, b_l and x.
Run:  python tests/golden/make_golden.py
"""
import json
from fractions import Fraction as Fr
from pathlib import Path

src = [1, 3, 1, 2, 4, 0, 2, 5, 3]
dst = [0, 1, 0, 2, 2, 3, 4, 4, 0]      # node 5 has no in-edge; 1->0 twice; 2->2 self loop
n, F, O = 6, 3, 2
x = [[Fr(i * 3 + f - 4, 2) for f in range(F)] for i in range(n)]
Wl = [[Fr(1), Fr(-2), Fr(1, 2)], [Fr(3, 4), Fr(0), Fr(-1)]]
Wr = [[Fr(-1, 2), Fr(1), Fr(2)], [Fr(1), Fr(1, 4), Fr(-3, 2)]]
bl = [Fr(1, 8), Fr(-5, 4)]
c = [[Fr((i + 1) * (o + 2) % 5 - 2, 4) for o in range(O)] for i in range(n)]   # upstream gradient

cnt = [0] * n
ssum = [[Fr(0)] * F for _ in range(n)]
for s, d in zip(src, dst):
    cnt[d] += 1
    for f in range(F):
        ssum[d][f] += x[s][f]
mean = [[ssum[i][f] / max(cnt[i], 1) for f in range(F)] for i in range(n)]
out = [[sum(Wl[o][f] * mean[i][f] + Wr[o][f] * x[i][f] for f in range(F)) + bl[o] for o in range(O)] for i in range(n)]

dWl = [[sum(c[i][o] * mean[i][f] for i in range(n)) for f in range(F)] for o in range(O)]
dWr = [[sum(c[i][o] * x[i][f] for i in range(n)) for f in range(F)] for o in range(O)]
db = [sum(c[i][o] for i in range(n)) for o in range(O)]
dmean = [[sum(c[i][o] * Wl[o][f] for o in range(O)) for f in range(F)] for i in range(n)]
dx = [[sum(c[i][o] * Wr[o][f] for o in range(O)) for f in range(F)] for i in range(n)]
for s, d in zip(src, dst):
    for f in range(F):
        dx[s][f] += dmean[d][f] / max(cnt[d], 1)

fl = lambda m: [[float(v) for v in r] for r in m]
gold = dict(src=src, dst=dst, n=n, x=fl(x), w_l=fl(Wl), w_r=fl(Wr), b_l=[float(v) for v in bl], grad_out=fl(c),
            mean=fl(mean), out=fl(out), d_w_l=fl(dWl), d_w_r=fl(dWr), d_b_l=[float(v) for v in db], d_x=fl(dx))
Path(__file__).with_name("sage_known_answer.json").write_text(json.dumps(gold))
print("wrote", Path(__file__).with_name("sage_known_answer.json"))
