import sys, os
sys.path.insert(0, os.getcwd())
import torch, cdlnet_video_b200 as cb
d = torch.device("cuda", 0)
K, M = 2, 169
plan = cb.Plan(3, 1, 1, M, K, (8, 32, 32), (7, 7, 7), 2, precision="tf32")
g = torch.Generator().manual_seed(0)
A = torch.randn(K, M, 1, 7, 7, 7, generator=g) * 0.05
plan.set_weights(A.to(d), A.to(d), torch.rand(K, 2, M, device=d) * 0.01)
y = torch.rand(1, 1, 8, 32, 32, device=d)
yp, _, mean = plan.preprocess(y)
z = plan.new_code()
c = torch.full((1,), 0.1, device=d)
plan.analysis_step(0, yp, z, c, first=True)
torch.cuda.synchronize()
print("ok", float(plan.export_code(z).abs().sum()))
