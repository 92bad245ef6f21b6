import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gym_kmanip_b200.vector_env import KManipVectorEnv
env = KManipVectorEnv("KManipSoloArmVision", 1024, seed=0)
env.reset()
for _ in range(4):
    env.step(env.sample_actions())
torch.cuda.synchronize()
print("ok")
