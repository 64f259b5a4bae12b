import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bbdm_b200.unet as U
from model.BrownianBridge.BrownianBridgeModel import BrownianBridgeModel
from torch.profiler import profile, ProfilerActivity
cfg = bench.CONFIGS["cfg3"]
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
net = BrownianBridgeModel(bench.namespace(cfg["unet"], 200)).train(); bench.init_weights(net.denoise_fn); net = net.cuda()
x = bench.synth((32, 3, 64, 64), 1).cuda(); y = bench.synth((32, 3, 64, 64), 2).cuda()
opt = torch.optim.Adam(net.get_parameters(), lr=1e-4)
def step():
    opt.zero_grad(set_to_none=True); loss, _ = net(x, y); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
