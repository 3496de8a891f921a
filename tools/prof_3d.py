"""Two forwards of NCSN3DShallow on 512 patches of 8x8x24 (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import parity_cases as C
from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsn3d import NCSN3DShallow
dev = torch.device("cuda")
cfg_T = C.make_config("CINE127", 128, 24, 400, 40.0, device="cuda")
cfg_T.data.channels, cfg_T.data.channels_3d = 64, 1
torch.manual_seed(1)
net_T = NCSN3DShallow(cfg_T).to(dev).eval()
vol = torch.rand(512, 8, 24, 8, device=dev); gv = torch.empty_like(vol); lab = torch.zeros(512, dtype=torch.long, device=dev)
for _ in range(2):
    net_T.forward_into(vol, lab, gv)
torch.cuda.synchronize()
print("ok")
