"""print_args / save_args_to_yaml used by the harness (reference utils/project_utils.py:30-109)."""
import os

import yaml


def print_args(args):
    d = vars(args)
    width = max(len(k) for k in d) if d else 0
    print("=" * 80)
    print("CONFIGURATION PARAMETERS:")
    print("=" * 80)
    for k in sorted(d):
        print(f"  {k.ljust(width)} : {d[k]}")
    print("=" * 80)
    print(f"Total number of cfg parameters: {len(d)}")
    print("=" * 80)


def save_args_to_yaml(args, output_path):
    """rank-0 only, like the reference (:47-109)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            return
    except Exception:  # noqa: BLE001
        pass
    os.makedirs(output_path, exist_ok=True)
    with open(os.path.join(output_path, "args.yaml"), "w") as f:
        yaml.safe_dump({k: (v if isinstance(v, (int, float, str, bool, list, type(None))) else str(v))
                        for k, v in vars(args).items()}, f, sort_keys=True)
