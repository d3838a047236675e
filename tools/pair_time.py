"""Timing of the fused conv pair at 1080p (see tools/pair_check.py).  python tools/pair_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.pair_check import time_pair

if __name__ == '__main__':
    time_pair()
