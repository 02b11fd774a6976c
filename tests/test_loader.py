"""Input row N2, host half: the WFDB format-16 batch loader (reader thread -> pinned ring -> device decode).
CPU tests cover validity checks, batching / shuffling / ragged batches and error propagation of the reader stage;
the GPU test runs the whole pipeline against the numpy oracle and through the reference-shaped eval loop."""
import os

import numpy as np
import pytest
import torch

from oracle import wfdb16_oracle as W

LEADS = ["I", "II", "III", "AVR", "AVL", "AVF", "V1", "V2", "V3", "V4", "V5", "V6"]


def write_record(base, rel, T, seed, gain=1000.0, n_leads=12, truncate=0):
    rng = np.random.default_rng(seed)
    frames = (rng.standard_normal((T, n_leads)) * 300 + rng.integers(-200, 200, size=(1, n_leads))).astype("<i2")
    path = os.path.join(base, rel)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    name = os.path.basename(rel)
    with open(path + ".hea", "w") as f:
        f.write(f"{name} {n_leads} 100 {T}\n")
        for l in range(n_leads):
            f.write(f"{name}.dat 16 {gain}(0)/mV 16 0 {int(frames[0, l])} 0 0 {LEADS[l % 12]}\n")
    raw = frames.tobytes()
    with open(path + ".dat", "wb") as f:
        f.write(raw[:len(raw) - truncate] if truncate else raw)
    return frames


@pytest.fixture()
def records(tmp_path):
    base = str(tmp_path)
    rels = [f"records100/00000/{i:05d}_lr" for i in range(11)]
    frames = [write_record(base, r, 200, seed=i) for i, r in enumerate(rels)]
    return base, rels, frames


def test_validate_records(records):
    from ptbxl_multimodal_b200.loader import validate_records
    base, rels, _ = records
    write_record(base, "records100/00000/bad_trunc", 200, seed=99, truncate=10)
    write_record(base, "records100/00000/bad_leads", 200, seed=98, n_leads=11)
    write_record(base, "records100/00000/no_dat", 200, seed=97)
    os.remove(os.path.join(base, "records100/00000/no_dat.dat"))
    with open(os.path.join(base, "records100/00000/garbage.hea"), "w") as f:
        f.write("garbage\n")
    mask = validate_records(base, rels + ["records100/00000/bad_trunc", "records100/00000/bad_leads",
                                          "records100/00000/no_dat", "records100/00000/garbage", "records100/00000/absent"])
    assert mask.tolist() == [True] * 11 + [False] * 5


def test_reader_stage_batches_in_order_with_ragged_tail(records):
    from ptbxl_multimodal_b200.loader import Wfdb16BatchLoader
    base, rels, frames = records
    y = np.arange(11 * 5, dtype=np.float32).reshape(11, 5)
    ld = Wfdb16BatchLoader(base, rels, y, batch_size=4, device="cpu")
    assert len(ld) == 3 and len(ld.dataset) == 11 and ld.T == 200
    seen = []
    for buf, idx, _ in ld.iter_host_batches():
        for j, rec in enumerate(idx):
            assert np.array_equal(buf[j], frames[rec])
        seen += [int(i) for i in idx]
    assert seen == list(range(11))
    ld2 = Wfdb16BatchLoader(base, rels, y, batch_size=4, device="cpu", drop_last=True)
    assert len(ld2) == 2 and sum(len(idx) for _, idx, _ in ld2.iter_host_batches()) == 8
    with pytest.raises(Exception):
        iter(ld).__next__()                         # decoding has no CPU fallback


def test_reader_stage_shuffle_is_seeded_and_changes_per_epoch(records):
    from ptbxl_multimodal_b200.loader import Wfdb16BatchLoader
    base, rels, frames = records
    y = np.zeros((11, 5), dtype=np.float32)

    def epoch(ld):
        out = []
        for buf, idx, _ in ld.iter_host_batches():
            for j, rec in enumerate(idx):
                assert np.array_equal(buf[j], frames[rec])
            out += [int(i) for i in idx]
        return out
    a = Wfdb16BatchLoader(base, rels, y, 3, "cpu", shuffle=True, seed=5, depth=3)
    b = Wfdb16BatchLoader(base, rels, y, 3, "cpu", shuffle=True, seed=5)
    e0, e1 = epoch(a), epoch(a)
    assert sorted(e0) == list(range(11)) and sorted(e1) == list(range(11))
    assert e0 != e1 and e0 != list(range(11))
    assert epoch(b) == e0


def test_reader_errors_surface_in_the_consumer(records):
    from ptbxl_multimodal_b200.loader import Wfdb16BatchLoader
    base, rels, _ = records
    write_record(base, "records100/00000/short", 200, seed=50, truncate=24)
    write_record(base, "records100/00000/gain", 200, seed=51, gain=500.0)
    y = np.zeros((12, 5), dtype=np.float32)
    with pytest.raises(RuntimeError, match="Failed to read record"):
        list(Wfdb16BatchLoader(base, rels + ["records100/00000/short"], y, 4, "cpu").iter_host_batches())
    with pytest.raises(RuntimeError, match="differs"):
        list(Wfdb16BatchLoader(base, rels + ["records100/00000/gain"], y, 4, "cpu").iter_host_batches())
    with pytest.raises(Exception):
        Wfdb16BatchLoader(base, rels, y[:3], 4, "cpu")


@pytest.mark.gpu
def test_loader_end_to_end_matches_oracle_and_feeds_the_eval_loops(tmp_path):
    import ptbxl_multimodal_b200 as P
    base = str(tmp_path)
    rels = [f"records100/00000/{i:05d}_lr" for i in range(21)]
    frames = [write_record(base, r, 1000, seed=i) for i, r in enumerate(rels)]
    rng = np.random.default_rng(3)
    y = (rng.random((21, 5)) < 0.3).astype(np.float32)
    demo = rng.random((21, 5)).astype(np.float32)
    ld = P.Wfdb16BatchLoader(base, rels, y, batch_size=8, device="cuda:0", demo=demo)
    h = P.loader.read_header(os.path.join(base, rels[0]))
    n = 0
    for x, d, yy in ld:
        b = x.shape[0]
        ref = np.stack([W.load_and_normalize(frames[n + j].tobytes(), h.gains, h.baselines) for j in range(b)])
        assert x.shape == (b, 12, 1000) and x.is_cuda
        assert np.abs(x.cpu().numpy() - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
        assert np.array_equal(yy.cpu().numpy(), y[n:n + b]) and np.array_equal(d.cpu().numpy(), demo[n:n + b])
        n += b
    assert n == 21
    # the reference-shaped loops run on it unchanged
    torch.manual_seed(0)
    model = P.ECGCNN(12, 256, 5).to("cuda:0")
    ld2 = P.Wfdb16BatchLoader(base, rels, y, batch_size=8, device="cuda:0", shuffle=True)
    opt = P.FusedAdamW(model.parameters(), lr=1e-3)
    loss = P.train_one_epoch(model, ld2, opt, "cuda:0")
    m = P.eval_one_epoch(model, ld2, "cuda:0", engine=P.InferStep(model, 8, 1000))
    assert np.isfinite(loss) and np.isfinite(m["bce_loss"]) and "auroc_macro" in m


def test_rank_shards_are_disjoint_equal_and_cover_the_epoch(records):
    from ptbxl_multimodal_b200.loader import Wfdb16BatchLoader
    base, rels, frames = records
    y = np.zeros((11, 5), dtype=np.float32)
    seen = []
    for r in range(3):
        ld = Wfdb16BatchLoader(base, rels, y, 2, "cpu", shuffle=True, seed=9, rank=r, world_size=3)
        assert len(ld.dataset) == 3 and len(ld) == 2
        mine = []
        for buf, idx, _ in ld.iter_host_batches():
            for j, rec in enumerate(idx):
                assert np.array_equal(buf[j], frames[rec])
            mine += [int(i) for i in idx]
        assert len(mine) == 3
        seen.append(mine)
    flat = sum(seen, [])
    assert len(set(flat)) == 9                                   # disjoint; 11 // 3 * 3 records per epoch
    with pytest.raises(Exception):
        Wfdb16BatchLoader(base, rels, y, 2, "cpu", rank=3, world_size=3)
