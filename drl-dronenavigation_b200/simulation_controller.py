"""Entry point with the reference's flow (Sol/Model/simulation_controller.py:77-137):
parse flags -> seed -> Track(circle r=1, 6 points, h=1) -> PBDroneSimulator -> run by --run_type."""
import random

import numpy as np
import torch

from . import waypoints as Waypoints
from .argparser import parse_args
from .simulator import PBDroneSimulator
from .waypoints import Track


def main(argv=None):
    args = parse_args(argv)
    print("Initial parsed arguments:")
    print(args)
    random.seed(args.seed); np.random.seed(args.seed); torch.manual_seed(args.seed)
    track = Track(Waypoints.circle(radius=1, num_points=6, height=1), circle=True)
    sim = PBDroneSimulator(args, track, target_factor=0)
    if args.run_type in ("full", "cont"):
        sim.run_full_training()
    elif args.run_type == "test":
        sim.run_test()
    elif args.run_type == "saved":
        print(sim.test_saved())
    elif args.run_type == "learning":
        sim.test_learning()
    else:
        raise NotImplementedError(f"--run_type {args.run_type}")


if __name__ == "__main__":
    main()
