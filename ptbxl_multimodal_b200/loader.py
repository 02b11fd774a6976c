"""Host half of the input row: a batch loader that feeds the step straight from WFDB format-16
records on disk, replacing the reference's per-item ``wfdb.rdsamp`` Dataset + DataLoader
(src/datasets/ptbxl.py:14-50,129-142; ptbxl_af.py, ptbxl_ecg_multimodal.py analogous) and its
construct-time validity scan that decodes every record once (ptbxl.py:53-71,101-106).

B200-first structure: the host never touches samples.  A reader thread ``readinto``s the raw .dat bytes of a
whole batch into a pinned (B, T, 12) int16 buffer (ring of ``depth`` buffers); the batch goes H2D on a copy
stream while the previous batch computes; decode + transpose + per-lead z-score run as ONE kernel on the device
(``ecgb200_wfdb16_zscore_f32``) on the consumer's stream.  Per window that is 120 KB (12x5000 int16) over PCIe
instead of 240 KB of float32, and zero host arithmetic.  Validity is checked from the header and the file size
only (``validate_records``): no record is decoded twice.

Iterating yields what the reference's DataLoaders yield -- ``(x, y)`` or ``(x_ecg, x_demo, y)`` -- with the
tensors already on the device, so ``train_one_epoch`` / ``eval_one_epoch`` (and the ``_demo`` variants) run
unchanged (their ``.to(device)`` calls become no-ops); ``loader.dataset`` has the ``__len__`` they use."""
from __future__ import annotations

import os
import queue
import threading
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import EcgB200Error
from .wfdb16 import Wfdb16Header, parse_header, decode_batch


def read_header(rec_path: str) -> Wfdb16Header:
    with open(rec_path + ".hea", "r") as f:
        return parse_header(f.read())


def record_is_valid(rec_path: str, n_leads: int = 12) -> bool:
    """Header-and-size check standing in for the reference's `_is_valid_ecg` (ptbxl.py:53-71), which decodes the
    whole record: .hea and .dat exist, the header parses as `n_leads` format-16 signals in one .dat file, and the
    .dat holds exactly n_samples frames."""
    try:
        h = read_header(rec_path)
        if h.n_sig != n_leads or len(set(h.dat_files)) != 1 or h.n_samples <= 0:
            return False
        return os.path.getsize(rec_path + ".dat") == h.n_samples * n_leads * 2
    except (OSError, ValueError, EcgB200Error):
        return False


def validate_records(base_dir: str, rel_paths: Sequence[str], n_leads: int = 12) -> np.ndarray:
    """Boolean mask over `rel_paths` (record paths without extension, relative to base_dir)."""
    return np.array([record_is_valid(os.path.join(base_dir, r), n_leads) for r in rel_paths], dtype=bool)


class _Records:
    """What the reference loops ask of ``loader.dataset``: a length."""

    def __init__(self, n: int):
        self.n = n

    def __len__(self) -> int:
        return self.n


class Wfdb16BatchLoader:
    def __init__(self, base_dir: str, rel_paths: Sequence[str], labels, batch_size: int, device,
                 demo=None, normalize: str = "per_lead", shuffle: bool = False, seed: int = 0,
                 drop_last: bool = False, depth: int = 2, n_leads: int = 12, rank: int = 0, world_size: int = 1):
        if len(rel_paths) == 0:
            raise EcgB200Error("no records")
        if batch_size <= 0 or depth < 2:
            raise EcgB200Error("batch_size must be positive and depth >= 2 (double buffering)")
        self.paths: List[str] = [os.path.join(base_dir, r) for r in rel_paths]
        self.device = torch.device(device)
        self.B, self.depth, self.n_leads = int(batch_size), int(depth), int(n_leads)
        self.shuffle, self.seed, self.drop_last = bool(shuffle), int(seed), bool(drop_last)
        self.normalize = normalize == "per_lead"
        # data parallel: every rank builds the same loader and reads its own interleaved share of each epoch's order
        # (equal counts on all ranks, so that the collective steps line up; the remainder is dropped)
        if world_size < 1 or not (0 <= rank < world_size):
            raise EcgB200Error("need 0 <= rank < world_size")
        self.rank, self.world = int(rank), int(world_size)
        self.epoch = 0
        h = read_header(self.paths[0])
        if h.n_sig != n_leads:
            raise EcgB200Error(f"Invalid lead count for {self.paths[0]}: {h.n_sig}, expected {n_leads}.")
        self.T, self.gains, self.baselines = h.n_samples, list(h.gains), list(h.baselines)
        self.frame_bytes = self.T * n_leads * 2
        labels = torch.as_tensor(labels).float()
        if labels.shape[0] != len(self.paths):
            raise EcgB200Error("one label row per record")
        self.cuda = self.device.type == "cuda"
        self.labels = labels.to(self.device) if self.cuda else labels
        self.demo = None
        if demo is not None:
            demo = torch.as_tensor(demo).float()
            if demo.shape[0] != len(self.paths):
                raise EcgB200Error("one demographic row per record")
            self.demo = demo.to(self.device) if self.cuda else demo
        self.dataset = _Records(len(self.paths) // self.world)      # what this rank sees per epoch
        self._hdr_ok = np.zeros(len(self.paths), dtype=bool)      # headers are checked once, not once per epoch

    def __len__(self) -> int:
        n = len(self.paths) // self.world
        return n // self.B if self.drop_last else (n + self.B - 1) // self.B

    def _batches(self) -> List[np.ndarray]:
        n = len(self.paths)
        order = np.random.default_rng(self.seed + self.epoch).permutation(n) if self.shuffle else np.arange(n)
        if self.world > 1:
            per = n // self.world
            order = order[:per * self.world][self.rank::self.world]
            n = per
        out = [order[i:i + self.B] for i in range(0, n, self.B)]
        if self.drop_last and out and len(out[-1]) < self.B:
            out.pop()
        return out

    # ------------------------------------------------------------------ host stage (reader thread)
    def _read_into(self, rec: int, dst: np.ndarray) -> None:
        """Raw bytes of record `rec` -> dst (T, n_leads) int16, after checking that its header agrees with the
        batch-wide gain / baseline the device decode uses."""
        path = self.paths[rec]
        if not self._hdr_ok[rec]:
            try:
                h = read_header(path)
            except (OSError, ValueError, EcgB200Error) as e:
                raise RuntimeError(f"Failed to read record {path}: {e}")
            if h.n_sig != self.n_leads:
                raise RuntimeError(f"Invalid lead count for {path}: {h.n_sig}, expected {self.n_leads}.")
            if h.n_samples != self.T or h.gains != self.gains or h.baselines != self.baselines:
                raise RuntimeError(f"Failed to read record {path}: header (n_samples/gain/baseline) differs from "
                                   f"the first record's; decode such records in a loader of their own")
            self._hdr_ok[rec] = True
        with open(path + ".dat", "rb") as f:
            got = f.readinto(memoryview(dst.reshape(-1).view(np.uint8)))
            if got != self.frame_bytes or f.read(1):
                raise RuntimeError(f"Failed to read record {path}: .dat holds {got}{'+' if got == self.frame_bytes else ''} "
                                   f"bytes, the header promises {self.frame_bytes}")

    def iter_host_batches(self, buffers: Optional[List[np.ndarray]] = None) -> Iterator[Tuple[np.ndarray, np.ndarray, int]]:
        """Reader-thread pipeline only: yields (frames (B, T, n_leads) int16 buffer, record indices, slot).  The
        buffer belongs to the loader again once the consumer asks for the next item `depth - 1` items later."""
        if buffers is None:
            buffers = [np.empty((self.B, self.T, self.n_leads), dtype=np.int16) for _ in range(self.depth)]
        batches = self._batches()
        self.epoch += 1
        free_q: "queue.Queue[int]" = queue.Queue()
        ready_q: "queue.Queue" = queue.Queue()
        for s in range(len(buffers)):
            free_q.put(s)
        stop = threading.Event()

        def reader():
            try:
                for idx in batches:
                    s = free_q.get()
                    if stop.is_set():
                        return
                    for j, rec in enumerate(idx):
                        self._read_into(int(rec), buffers[s][j])
                    ready_q.put((s, idx))
                ready_q.put(None)
            except BaseException as e:             # surfaces in the consumer
                ready_q.put(e)

        th = threading.Thread(target=reader, daemon=True, name="ecgb200-wfdb16-reader")
        th.start()
        held: List[int] = []
        try:
            while True:
                item = ready_q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                s, idx = item
                held.append(s)
                if len(held) > len(buffers) - 1:   # the oldest outstanding buffer goes back to the reader
                    free_q.put(held.pop(0))
                yield buffers[s], idx, s
        finally:
            stop.set()
            free_q.put(0)                          # unblock a reader waiting for a buffer

    # ------------------------------------------------------------------ device stage
    def __iter__(self):
        if not self.cuda:
            raise EcgB200Error("Wfdb16BatchLoader decodes on the device: it needs a CUDA device (no CPU fallback)")
        dev = self.device
        pinned = [torch.empty((self.B, self.T, self.n_leads), dtype=torch.int16).pin_memory() for _ in range(self.depth)]
        host = [p.numpy() for p in pinned]
        devbuf = [torch.empty((self.B, self.T, self.n_leads), dtype=torch.int16, device=dev) for _ in range(self.depth)]
        copied = [torch.cuda.Event() for _ in range(self.depth)]
        decoded = [torch.cuda.Event() for _ in range(self.depth)]
        copy_stream = torch.cuda.Stream(device=dev)
        src = self.iter_host_batches(host)

        def stage():
            """Next host batch -> device frames on the copy stream (returns None at the end)."""
            try:
                _, idx, s = next(src)
            except StopIteration:
                return None
            copy_stream.wait_event(decoded[s])            # the decode that last read devbuf[s] has finished
            with torch.cuda.stream(copy_stream):
                devbuf[s][:len(idx)].copy_(pinned[s][:len(idx)], non_blocking=True)
                copied[s].record(copy_stream)
            return s, idx

        nxt = stage()
        while nxt is not None:
            s, idx = nxt
            # the host buffer must not be refilled before its H2D has left it: the reader gets it back only after
            # the NEXT item is requested, by which time this wait has long passed
            copied[s].synchronize()
            nxt = stage()                                 # batch i+1 streams in under batch i's compute
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(copied[s])
            x = decode_batch(devbuf[s][:len(idx)], self.gains, self.baselines, normalize=self.normalize)
            decoded[s].record(cur)
            sel = torch.as_tensor(idx, device=dev)
            y = self.labels.index_select(0, sel)
            if self.demo is not None:
                yield x, self.demo.index_select(0, sel), y
            else:
                yield x, y
