"""Small case for compute-sanitizer: every mapping, a few env steps with autoreset, reset, contacts, site poses."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from gym_kmanip_b200.batch_sim import BatchSim
for env in ("KManipSoloArm", "KManipDualArm"):
    for lanes in (32, 1, 2):
        s = BatchSim(env, 40, dtype="float32", seed=1, max_episode_steps=3)
        s.configure(lanes, 0)
        s.reset()
        gen = torch.Generator(device="cuda").manual_seed(0)
        for t in range(5):
            s.step(torch.rand(40, s.act_dim, device="cuda", generator=gen) * 2 - 1, autoreset=True)
        s.contacts(); s.site_poses(); s.get_state()
        torch.cuda.synchronize()
        s.close()
        print(env, lanes, "ok", flush=True)
# cost-ordered walk (counting sort + dynamic tile fetch) on small tiles, and the exact-parity IK (work arrays in dynamic
# shared memory behind the env records)
for env, lanes, epb, kw in (("KManipSoloArmQPos", 32, 1, {}), ("KManipSoloArm", 32, 2, dict(ik_mode=1)), ("KManipDualArm", 32, 3, dict(ik_mode=1))):
    s = BatchSim(env, 333, dtype="float32", seed=1, max_episode_steps=3, **kw)
    s.configure(lanes, epb)
    s.set_env_ordering(1)
    s.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(4):
        s.step(torch.rand(333, s.act_dim, device="cuda", generator=gen) * 2 - 1, autoreset=True)
    torch.cuda.synchronize()
    s.close()
    print(env, lanes, epb, kw, "ordered walk ok", flush=True)
# camera observations: setup + pixel kernels, aligned (640 x 480) and unaligned (60 x 40) store paths, partial tiles
for env, cams in (("KManipSoloArmVision", ("head", "grip_r")), ("KManipTorsoVision", ("top", "grip_l"))):
    s = BatchSim(env, 3, dtype="float32", seed=1)
    s.reset()
    for c in cams:
        img = s.render(c)
        assert img.any()
    s.render_records()
    torch.cuda.synchronize()
    s.close()
    print(env, "render ok", flush=True)
