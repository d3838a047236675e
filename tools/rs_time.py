"""Per-layer timing of the 48->48 3x3 conv at 1080p on the row-streaming (rs) and tile (tc) kernels.
    python tools/rs_time.py [label ...]      labels: plain silu mish gate"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tools.rs_check import time_layers

if __name__ == '__main__':
    time_layers()
