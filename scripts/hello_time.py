import sys; sys.path.insert(0,'/root/repo')
from scripts.quick_time import run
run("hello", 65536, 8, 5, max_episode_steps=100)
run("hello", 65536, 32, 5, max_episode_steps=100)
