import os, sys, time, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_l2o_b200 import hrnn_train as ht
DEV = "cuda:0"
sizes = [864, 32, 25600, 32, 327680, 10]
eng = ht._Engine(sizes, DEV)
N, nt = eng.N, eng.nt
theta = ht._init_theta(0).to(DEV)
planes = torch.rand(21, N, device=DEV) * 0.5 + 0.1
g = torch.randn(N, device=DEV) * 0.1
bias0 = torch.zeros(nt, 32, device=DEV); mean = torch.zeros(1, device=DEV); zf = torch.zeros(nt, 4, dtype=torch.int32, device=DEV)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("coord_forward", lambda: eng.coord_forward(theta, planes, bias0, mean, g, zf)),
                 ("coord_backward", lambda: eng.coord_backward(theta, planes, bias0, mean, g, zf, torch.ones(21, N, device=DEV), torch.ones(N, device=DEV), torch.ones(nt, 24, device=DEV)))):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "%.3f ms per call at N=%d" % (e0.elapsed_time(e1) / 5, N))
