"""One traced launch of the fused pair at 1080p: RSB_PAIR_TRACE=<file> python tools/pair_trace.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.engine import native as N
from tools.pair_check import DEV, build

plan, _ = build(N.ACT_SILU, True)
x = torch.randn(1, 48, 1080, 1920).to(DEV, torch.bfloat16)
plan.forward(x)
plan.forward(x, ops=(1, 3))
torch.cuda.synchronize()
