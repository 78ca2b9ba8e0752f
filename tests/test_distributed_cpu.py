"""CPU (gloo, world_size 2): pair sharding, the single all-gather and the chain composition of the multi-GPU path."""

import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from spateo_release_b200.alignment.distributed import gather_transformations, morpho_align_chain_sharded, shard_pairs
from spateo_release_b200.alignment.morpho_alignment import compose_transformations, morpho_align_apply_transformation
from spateo_release_b200.anndata_lite import AnnDataLite


def _rot(th):
    return np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])


def _fake_pair(modelA, modelB, spatial_key="spatial", **kw):
    """Stand-in for the GPU solver: the exact similarity between two slices that are rigid copies of each other."""
    a, b = modelA.uns["pose"], modelB.uns["pose"]
    # raw_B = base R_b^T + t_b ; raw_A = base R_a^T + t_a  ->  aligned_B = raw_B R^T + t with R = R_a R_b^T
    R = _rot(a[0]) @ _rot(b[0]).T
    t = np.array(a[1:]) - np.array(b[1:]) @ R.T
    return {"Rotation": R, "Translation": t}


def _models(n=7):
    rng = np.random.default_rng(0)
    base = rng.uniform(0, 10, size=(50, 2))
    out = []
    for k in range(n):
        pose = (0.2 * k, 1.0 * k, -0.5 * k)
        raw = base @ _rot(pose[0]).T + np.array(pose[1:])
        out.append(AnnDataLite(np.zeros((50, 3), dtype=np.float32), obsm={"spatial": raw}, uns={"pose": pose}))
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    models = _models()
    models, tr = morpho_align_chain_sharded(models, pair_fn=_fake_pair)
    q.put((rank, [np.asarray(m.obsm["align_spatial"]) for m in models], tr))
    dist.destroy_process_group()


def test_shard_pairs_round_robin():
    assert shard_pairs(15, 0, 8) == [0, 8] and shard_pairs(15, 7, 8) == [7]
    assert sorted(sum((shard_pairs(63, r, 8) for r in range(8)), [])) == list(range(63))


def test_chain_sharded_world2_matches_serial():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # serial reference: the driver functions of the single-process path
    models = _models()
    tr = [_fake_pair(models[i], models[i + 1]) for i in range(len(models) - 1)]
    serial = morpho_align_apply_transformation(models, transformation=tr)
    want = [np.asarray(m.obsm["align_spatial"]) for m in serial]
    for rank, got, tr_g in results:
        assert len(tr_g) == len(tr)
        for a, b in zip(got, want):
            assert np.allclose(a, b, atol=1e-10)
    # every slice lands on slice 0's frame (all slices are rigid copies of one base)
    for a in want[1:]:
        assert np.allclose(a, want[0], atol=1e-9)


def test_gather_single_process():
    tr = {0: {"Rotation": _rot(0.3), "Translation": np.array([1.0, 2.0])}, 1: {"Rotation": _rot(-0.1), "Translation": np.zeros(2)}}
    out = gather_transformations(tr, 2)
    assert np.allclose(out[0]["Rotation"], _rot(0.3)) and np.allclose(out[1]["Translation"], 0)
    (R1, t1), (R2, t2) = compose_transformations(out)
    assert np.allclose(R2, _rot(0.3) @ _rot(-0.1))


def test_transformation_resume_recomputes_only_the_last_checkpointed_pair(tmp_path, monkeypatch):
    """``resume=True`` (morpho_alignment.py:166-179): restart at the highest checkpointed pair, return exactly
    len(models) - 1 plain dicts equal to an uninterrupted run, and feed morpho_align_apply_transformation unchanged."""
    from spateo_release_b200.alignment import morpho_alignment as ma

    calls = []

    def counting_pair(modelA, modelB, spatial_key="spatial", **kw):
        calls.append((modelA.uns["pose"], modelB.uns["pose"]))
        return _fake_pair(modelA, modelB, spatial_key=spatial_key)

    monkeypatch.setattr(ma, "pair_transformation", counting_pair)
    models = _models(5)
    path = str(tmp_path / "tr")
    full = ma.morpho_align_transformation(models, save_transformation=True, transformation_path=path, verbose=False)
    assert len(full) == 4 and len(calls) == 4
    os.remove(os.path.join(path, "transformation_3.npy"))  # an interrupted run: pairs 0..2 are on disk
    calls.clear()
    resumed = ma.morpho_align_transformation(models, save_transformation=True, transformation_path=path, resume=True,
                                             verbose=False)
    assert len(resumed) == 4 and all(isinstance(t, dict) for t in resumed)
    assert len(calls) == 2, "pairs 2 (highest checkpoint, recomputed like the reference) and 3 only"
    for a, b in zip(full, resumed):
        assert np.allclose(a["Rotation"], b["Rotation"]) and np.allclose(a["Translation"], b["Translation"])
    assert sorted(os.listdir(path)) == [f"transformation_{i}.npy" for i in range(4)]
    placed = ma.morpho_align_apply_transformation(_models(5), transformation=resumed, verbose=False)
    ref = np.asarray(placed[0].obsm["align_spatial"])
    for m in placed[1:]:
        assert np.abs(np.asarray(m.obsm["align_spatial"]) - ref).max() < 1e-9
    # and from the checkpoints on disk (np.load(..., allow_pickle=True) objects)
    placed2 = ma.morpho_align_apply_transformation(_models(5), transformation=None, transformation_path=path, verbose=False)
    assert np.abs(np.asarray(placed2[4].obsm["align_spatial"]) - ref).max() < 1e-9
