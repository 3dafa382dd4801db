"""Developer probe: does PPO on the batched engine learn?  python tools/ppo_train_probe.py [env_id] [epochs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phoenix_drone_simulation_b200.ppo import PPO
env_id = sys.argv[1] if len(sys.argv) > 1 else 'DroneHoverSimpleEnv-v0'
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 30
alg = PPO(env_id, num_envs=4096, steps=64, epochs=epochs, seed=0)
alg.learn(verbose=True)
