import sys, os
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
from scripts.hello_time import run
for n in (1 << 17, 1 << 18, 3 << 16):
    run(n, 32, 5, max_episode_steps=100)
