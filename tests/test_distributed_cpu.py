"""The N>1 path on CPU: world_size-2 gloo processes shard a batch by bone and one mesh by plane
range with no data-path collective; the only exchange is the final gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from shoulder_b200 import meshio, sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inner_zs(z):
    lo, hi = z.min(), z.max()
    return np.linspace(hi - 0.2 * (hi - lo), lo + 0.2 * (hi - lo), 4)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), GLOO_SOCKET_IFNAME="lo")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    v, f = meshio.icosphere(2, 1.0, scale=(20.0, 30.0, 170.0))
    # --- by bone: 6 jittered bones over 2 ranks
    mine = sharding.shard_bones(6, rank, world)
    cents = []
    for b in mine:
        m = meshio.synthetic_bone(meshio.Mesh(v, f), int(b))
        zs = _inner_zs(m.vertices[:, 2])
        cents.append(oracle.OracleSlices(m.vertices, m.faces, zs, 16).centroids)
    payload = (mine, np.array(cents))
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)                 # the single final gather
    # --- by plane range on one mesh: concatenation must reproduce the unsharded sweep exactly
    zs = np.linspace(150.0, -150.0, 11)
    z_orig, h, (lo, hi) = sharding.plane_shard_heights(zs, rank, world)
    paths = oracle.section_multiplane(v, f, [0, 0, z_orig], [0, 0, 1], h)
    part = np.array([p.centroid for p in paths])
    parts = [None] * world
    dist.all_gather_object(parts, part)
    # --- block-cyclic deal of the same sweep (what bench.py uses for one large mesh): scatter by index
    z_orig2, h2, idx = sharding.plane_shard_heights_cyclic(zs, rank, world, block=2)
    paths2 = oracle.section_multiplane(v, f, [0, 0, z_orig2], [0, 0, 1], h2)
    cyc = [None] * world
    dist.all_gather_object(cyc, (idx, np.array([p.centroid for p in paths2])))
    if rank == 0:
        scattered = np.zeros((len(zs), 2))
        for i, c in cyc:
            scattered[i] = c
        out.put((gathered, np.concatenate(parts), scattered))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        if p.exitcode is None:
            p.kill()
        assert p.exitcode == 0, "gloo worker failed"
    gathered, cat, scattered = out.get()
    ids = np.concatenate([g[0] for g in gathered])
    assert np.array_equal(np.sort(ids), np.arange(6))
    v, f = meshio.icosphere(2, 1.0, scale=(20.0, 30.0, 170.0))
    for g in gathered:                                        # each shard equals the unsharded computation
        for b, c in zip(g[0], g[1]):
            m = meshio.synthetic_bone(meshio.Mesh(v, f), int(b))
            zs = _inner_zs(m.vertices[:, 2])
            assert np.array_equal(c, oracle.OracleSlices(m.vertices, m.faces, zs, 16).centroids)
    zs = np.linspace(150.0, -150.0, 11)
    whole = oracle.OracleSlices(v, f, zs, 16).centroids
    assert np.array_equal(cat, whole)
    assert np.array_equal(scattered, whole)
    for n, w in ((8192, 8), (100, 3)):
        got = np.sort(np.concatenate([sharding.shard_planes_cyclic(n, r, w) for r in range(w)]))
        assert np.array_equal(got, np.arange(n))
