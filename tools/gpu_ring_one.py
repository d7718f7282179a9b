import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops
dev = torch.device("cuda:0")
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n, t, h, w, cout = 48, 32, 56, 56, 64
x = (torch.randn(n, t, h, w, cin, device=dev) * 0.5).to(torch.bfloat16)
wt = torch.randn(cout, cin, 3, 1, 1, device=dev) / (cin * 3) ** 0.5
d = ops.conv_desc(n, t, h, w, cin, cout, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU)
wp = ops.pack_conv_weight(d, wt)
y = torch.empty(n, t, h, w, cout, device=dev, dtype=torch.bfloat16)
for _ in range(3): ops.conv3d_fwd(d, x, wp, out=y)
torch.cuda.synchronize()
