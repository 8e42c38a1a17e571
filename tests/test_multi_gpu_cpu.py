"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sample-index sharding + one film reduce.
The local renderer is stubbed with the CPU oracle; the code under test is mitsubaer_b200.distributed."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import make_field, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict
    from mitsubaer_b200 import distributed as mdist
    from oracle.oracle import Oracle, volume_desc
    orc = Oracle(np.float32)
    data, lo, hi = make_field("radial", 24)
    props = medium_props(stepsize=4e-2)
    omed = orc.medium_create(oracle_medium_desc(props, 0.9), orc.rif_create(volume_desc((24,) * 3, lo, hi), data))
    scene = scene_dict(24, 16, 6, rfilter="gaussian")
    film = torch.zeros(16, 24, 5)

    def render_local(begin, stride):
        f, st = orc.render(omed, oracle_render_desc(scene, sample_begin=begin, sample_stride=stride), nthreads=2)
        film.add_(torch.from_numpy(f))
        return st.as_dict()

    stats = mdist.render_sharded(render_local, film, dst=0)
    assert stats["samples"] == 24 * 16 * mdist.local_spp(6, rank, world)
    total = mdist.reduce_stats(stats)
    assert total["samples"] == 24 * 16 * 6
    # the widened path shards the same way: light paths + transient frames, each rank adding its share of the unit weight
    scene2 = dict(scene_dict(16, 16, 5, rfilter="box"), fov=30.0, envRadiance=0.0, transient=dict(minBound=2.0, maxBound=18.0, binWidth=2.0))
    film2 = torch.zeros(16, 16, 3 * 8 + 2)

    def render_light(begin, stride):
        f, st = orc.render(omed, oracle_render_desc(scene2, sample_begin=begin, sample_stride=stride, light_tracing=True, props=props), nthreads=2)
        film2.add_(torch.from_numpy(f))
        return st.as_dict()

    stats2 = mdist.reduce_stats(mdist.render_sharded(render_light, film2, dst=0))
    if rank == 0:
        full, st = orc.render(omed, oracle_render_desc(scene), nthreads=2)
        assert total["ray_steps"] == st.ray_steps
        assert np.allclose(film.numpy(), full, rtol=1e-5, atol=1e-6)
        full2, st2 = orc.render(omed, oracle_render_desc(scene2, light_tracing=True, props=props), nthreads=2)
        assert stats2["connections"] == st2.connections and stats2["samples"] == 16 * 16 * 5
        assert np.allclose(film2.numpy()[..., -1], 1.0) and np.allclose(film2.numpy(), full2, rtol=1e-5, atol=1e-6 * np.abs(full2).max() + 1e-7)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_render_gloo_world2(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_shard_arithmetic():
    sys.path.insert(0, ROOT)
    from mitsubaer_b200 import distributed as mdist
    for world in (1, 2, 3, 8):
        for spp in (1, 5, 8, 256):
            counts = [mdist.local_spp(spp, r, world) for r in range(world)]
            assert sum(counts) == spp
            seen = sorted(s for r in range(world) for s in range(mdist.shard(r, world)[0], spp, world))
            assert seen == list(range(spp))
