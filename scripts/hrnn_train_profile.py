import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_l2o_b200 import hrnn_train as ht
from open_l2o_b200.scale_problems import ConvNet
from torch.profiler import profile, ProfilerActivity
DEV = "cuda:0"
prob = ConvNet((3, 32, 32), 10, [(3, 3, 32), (5, 5, 32)])
params0 = [p.detach() for p in prob.init_tensors(seed=0, device=DEV)]
gen = torch.Generator().manual_seed(0)
data = torch.randn(128, 32, 32, 3, generator=gen).to(DEV)
labels = torch.nn.functional.one_hot(torch.randint(0, 10, (128,), generator=gen), 10).float().to(DEV)
tr = ht.MetaTrainer(prob.param_shapes, device=DEV, random_seed=0)
obj = lambda ps: prob.objective(ps, data, labels)
tr.train_step(obj, params0, 20)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.train_step(obj, params0, 20)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=10, max_name_column_width=60))
