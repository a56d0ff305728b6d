"""one ResBlock chain call for ncu (development aid): resblock_ncu.py [single|pair] B H W"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dagl_b200.resblock import ResBlock, resblocks_forward
mode = sys.argv[1]; B, H, W = (int(v) for v in sys.argv[2:5])
dev = torch.device("cuda"); torch.manual_seed(0)
blocks = [ResBlock(64).to(dev).eval() for _ in range(2)]
x = torch.randn(B, 64, H, W, device=dev)
with torch.no_grad():
    for _ in range(2): resblocks_forward(blocks, x, mode)
torch.cuda.synchronize()
