"""CPU-side checks of the oracle's image formation and of the multi-process (N>1) host logic."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, cornell_args


def test_oracle_reproduces_the_committed_golden_image(fb, oracle):
    gold = np.load(os.path.join(GOLDEN, "cornell_64_8spp.npz"))
    sc = fb.Scene(cornell_args(64, 4))
    fbuf = oracle.new_framebuffer(sc.view)
    ev = 0
    for i in range(8):
        ev += oracle.render_pass(sc.view, i, fbuf).shade_events
    assert ev == int(gold["shade_events"])
    assert np.array_equal(fbuf[5], gold["composited"])            # single-threaded or not: per-pixel work is independent
    assert np.array_equal(fbuf[4], gold["direct"])
    sc.close()


def test_oracle_direct_lighting_energy_partition(fb, oracle):
    """NEE-only, BSDF-only and MIS estimate the same direct lighting (toggles of src/renderers/pathtracer.h:206-247)."""
    means = {}
    for name, extra in (("mis", []), ("nee", ["-bsdf", "0"]), ("bsdf", ["-nee", "0"])):
        sc = fb.Scene(cornell_args(32, 1, extra))
        fbuf = oracle.new_framebuffer(sc.view)
        for i in range(96):
            oracle.render_pass(sc.view, i, fbuf)
        means[name] = float(fbuf[5][..., :3].mean())
        sc.close()
    assert abs(means["nee"] - means["mis"]) / means["mis"] < 0.04
    assert abs(means["bsdf"] - means["mis"]) / means["mis"] < 0.08


def test_oracle_pixel_subsets_are_independent(fb, oracle):
    sc = fb.Scene(cornell_args(32, 3))
    a = oracle.new_framebuffer(sc.view); b = oracle.new_framebuffer(sc.view)
    oracle.render_pass(sc.view, 0, a)
    px = np.arange(32 * 32, dtype=np.uint32)
    oracle.render_pass(sc.view, 0, b, pixels=px[::2])
    oracle.render_pass(sc.view, 0, b, pixels=px[1::2])
    assert np.array_equal(a, b)
    sc.close()


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, %(root)r)
    import torch, torch.distributed as dist
    import fermat_b200 as fb, oracle
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    args = ["-i", %(scene)r, "-r", "80", "48", "-bounces", "2", "-shard", str(rank), str(world)]
    sc = fb.Scene(args)
    fbuf = oracle.new_framebuffer(sc.view)
    own = sc.owned_pixels()
    ev = 0
    for i in range(3):
        ev += oracle.render_pass(sc.view, i, fbuf, pixels=own).shade_events
    img = torch.from_numpy(fbuf[5].copy())
    dist.reduce(img, dst=0, op=dist.ReduceOp.SUM)              # the single image reduce of the multi-GPU path
    evt = torch.tensor([ev], dtype=torch.int64)
    dist.all_reduce(evt)
    if rank == 0:
        np.savez(%(out)r, img=img.numpy(), events=evt.numpy())
    dist.destroy_process_group()
""")


def test_two_process_sharding_and_reduce_with_gloo(fb, oracle, tmp_path):
    """world_size 2 over gloo: each rank renders its tile shard, one reduce recombines the frame."""
    scene = os.path.join(GOLDEN, "cornellbox_jp.fbs")
    out = str(tmp_path / "reduced.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "scene": scene, "out": out})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), WORLD_SIZE="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e))
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = np.load(out)
    sc = fb.Scene(["-i", scene, "-r", "80", "48", "-bounces", "2"])
    fbuf = oracle.new_framebuffer(sc.view)
    ev = 0
    for i in range(3):
        ev += oracle.render_pass(sc.view, i, fbuf).shade_events
    assert np.array_equal(got["img"], fbuf[5])
    assert int(got["events"][0]) == ev
    sc.close()
