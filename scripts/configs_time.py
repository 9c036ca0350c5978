"""Development aid: step-kernel throughput of the BASELINE parity configs (1: Demo 1 at 65,536 envs; 2: Demo 2 and Demo 4
at 2^20 envs), same method as bench.py (fused 32-step launches, output ring larger than L2, CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.quick_time import run
run("demo1", 65536, 32, 50, max_episode_steps=100, track_returns=True)
run("demo1", 1 << 20, 32, 20, max_episode_steps=100, track_returns=True)
run("demo2", 1 << 20, 32, 20, max_episode_steps=100, track_returns=True)
run("demo4", 1 << 20, 32, 20, max_episode_steps=100, track_returns=True)
