"""World-size-2 run of the multi-GPU plumbing on CPU (gloo): the asset arena is broadcast once from
rank 0, every rank gets identical meshes and light maps (incl. precomputed IBL maps), the scene shards are disjoint and cover the batch."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from stillleben_b200 import dist as sdist
    from stillleben_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pool = synth.mesh_pool(3, nu=16, nv=8, tex_size=16) if rank == 0 else None
    lms = None
    if rank == 0:       # a light map whose IBL maps were "precomputed" on rank 0 (stand-in arrays of the real shapes: no GPU here)
        from stillleben_b200.desc import LightMapData
        eq, sun = synth.procedural_equirect(32, 16, seed=3)
        rng = np.random.RandomState(1)
        maps = (rng.rand(6, 8, 8, 4).astype(np.float32), rng.rand(6, 4, 4, 4).astype(np.float32),
                rng.rand(sum(6 * (16 >> m) ** 2 * 4 for m in range(5))).astype(np.float32), rng.rand(8, 8, 4).astype(np.float32))
        lms = [LightMapData(eq, [sun.tolist()], [[2.0, 1.9, 1.7]], maps), LightMapData(eq * 2, [], [])]
    pool, lms = sdist.broadcast_assets(pool, lms, 0)
    assert len(lms) == 2 and lms[0].maps is not None and lms[1].maps is None and lms[0].maps[0].shape == (6, 8, 8, 4)
    lm_digest = float(sum(np.float64(a.sum()) for a in lms[0].maps) + lms[0].equirect.sum() + lms[1].equirect.sum() + sum(lms[0].light_directions[0]))
    lo, hi = sdist.shard_range(37, rank, world)
    scenes = [synth.tabletop_scene(pool, 1000 + s, n_objects=4) for s in range(lo, hi)]
    digest = float(sum(np.float64(m.vertices["position"].sum()) + m.indices.sum() + sum(int(i.pixels.sum()) for i in m.images) for m in pool))
    q.put((rank, lo, hi, len(scenes), digest + lm_digest, float(scenes[0].objects[0].pose.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, n0, d0, _), (r1, lo1, hi1, n1, d1, _) = out
    assert (lo0, hi0, lo1, hi1) == (0, 19, 19, 37) and n0 + n1 == 37
    assert d0 == d1            # identical assets on both ranks after ONE broadcast
