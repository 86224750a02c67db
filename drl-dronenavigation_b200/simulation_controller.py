"""Entry point with the reference's flow (Sol/Model/simulation_controller.py:77-137):
parse flags -> seed -> Track(circle r=1, 6 points, h=1) -> PBDroneSimulator -> run by --run_type."""
import random

import numpy as np
import torch

from . import waypoints as Waypoints
from .argparser import parse_args
from .simulator import PBDroneSimulator
from .waypoints import Track


def _init_distributed():
    """Under torchrun: one process per GPU, NCCL for the learner's gradient all-reduce (the env shards need none)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank(), world


def _reject_out_of_scope(args):
    """Flags of the reference's parser that select subsystems outside the env-step path: accepted by the parser
    so that the reference's command lines parse, but selecting one fails loudly instead of silently running something else."""
    if getattr(args, "lib", "sb3") != "sb3":
        raise NotImplementedError(f"--lib {args.lib}: the Ray / TF-Agents / CleanRL launchers are out of scope; the torch-native learners replace --lib sb3")
    for flag, why in (("wandb", "wandb logging is out of scope (no network); use --tensorboard DIR"),
                      ("capture_video", "video capture needs the PyBullet renderer"), ("gui", "the PyBullet GUI is outside the CUDA hot path"),
                      ("vec_normalize", "the reference wraps a single gym env in VecNormalize here, which raises (PBDroneSimulator.py:186-189); "
                                        "NormalizeObservation is fused into the step kernel and on by default for every env (train, evaluation, VecEnv), "
                                        "--no_norm_obs turns it off, --norm_rew / --clip_rew select the reward wrappers"),
                      ("vec_check_nan", "the reference wraps a single gym env in VecCheckNan here, which raises (PBDroneSimulator.py:186-189)")):
        if getattr(args, flag, False):
            raise NotImplementedError(f"--{flag}: {why}")


def main(argv=None):
    args = parse_args(argv)
    _reject_out_of_scope(args)
    rank, world = _init_distributed()
    args.rank, args.world = rank, world
    if rank != 0:
        import builtins
        builtins.print = lambda *a, **k: None           # rank 0 logs
    print("Initial parsed arguments:")
    print(args)
    random.seed(args.seed); np.random.seed(args.seed); torch.manual_seed(args.seed)
    if getattr(args, "track", "circle") == "circle":     # the reference's main() (simulation_controller.py:96-100)
        track = Track(Waypoints.circle(radius=1, num_points=6, height=1), circle=True)
    else:                                                # its commented-out alternatives, selectable here
        track = Track(getattr(Waypoints, args.track)(), circle=False)
    sim = PBDroneSimulator(args, track, target_factor=0)
    if args.run_type in ("full", "cont"):
        log = print if rank == 0 else (lambda *a, **k: None)
        if getattr(args, "profile", False):              # simulation_controller.py:111-117
            from .profiler import Profiler
            with Profiler(print_fn=log):
                sim.run_full_training(log=log)
        else:
            sim.run_full_training(log=log)
    elif args.run_type == "test":
        sim.run_test()
    elif args.run_type == "saved":
        print(sim.test_saved())
    elif args.run_type == "learning":
        sim.test_learning()
    else:
        raise NotImplementedError(f"--run_type {args.run_type}")

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
