#!/usr/bin/env python
"""Multi-GPU parity of the frame gather, run under torchrun on N GPUs of one box:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py [--res W H]

Every rank renders its tile shard of the frame (scene replicated), the product's NCCL gather (fb200_context_gather_image) assembles the
frame on rank 0, and rank 0 compares it with its own UNSHARDED render of the same passes: they must be identical bit for bit
(disjoint tiles, per-pixel accumulation order unchanged). Prints one JSON line on rank 0 (also written to gpurun_out/ when present).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import fermat_b200 as fb
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, nargs=2, default=[1600, 900])
    ap.add_argument("--passes", type=int, default=4)
    ap.add_argument("--scene", default=os.path.join(ROOT, "scenes", "_cache", "bathroom2.fbs"))
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")            # only to hand the NCCL id around: the data path is the product's own communicator
    scene = args.scene if fb.scene_available(args.scene) else os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs")
    base = ["-i", scene, "-r", str(args.res[0]), str(args.res[1]), "-bounces", "8"]
    sc = fb.Scene(base + ["-shard", str(rank), str(world)])
    rc = fb.RenderingContext(sc, local)
    ids = [fb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    rc.comm_init(ids[0], rank, world)
    host = torch.zeros((args.res[1], args.res[0], 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    rc.clear()
    for i in range(args.passes):
        rc.render(i, sync=False)
        rc.gather_image(0, host.data_ptr() if (rank == 0 and i == args.passes - 1) else None)
    rc.synchronize()
    dist.barrier()
    if rank == 0:
        got = host.numpy().copy()
        dev = rc.gathered_tensor().cpu().numpy()
        full_sc = fb.Scene(base)
        full = fb.RenderingContext(full_sc, local)
        full.clear()
        for i in range(args.passes):
            full.render(i, sync=False)
        want = full.download()
        out = {"check": "frame gather vs unsharded render", "n_gpus": world, "res": args.res, "passes": args.passes, "scene": os.path.basename(scene),
               "bit_identical": bool(np.array_equal(got, want)), "device_copy_identical": bool(np.array_equal(dev, want)),
               "max_abs_diff": float(np.abs(got - want).max()), "mean": float(want[..., :3].mean())}
        line = json.dumps(out)
        print(line, flush=True)
        d = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(d):
            open(os.path.join(d, "multigpu_check_n%d.json" % world), "w").write(line + "\n")
        full.close(); full_sc.close()
    dist.barrier()
    rc.close(); sc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
