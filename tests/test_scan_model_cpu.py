"""Executable model of the grid-wide greedy NMS reduction (csrc/nms.cu: nms_scan_grid_kernel) checked against the C oracle's greedy NMS.

The kernel itself only runs on the GPU (tests/test_ops_gpu.py compares it with the oracle and the reference kernel bit for bit); what is
pinned here is the ALGORITHM: deciding 16 blocks (1024 boxes) at a time from the chunk-local words, carrying a chunk's contribution to
the next chunk through `s_next`, and letting the workers apply a chunk's kept rows to the global bitmap for all words from the chunk
AFTER the next one on, at any moment between the ticket's publication and the start of the first chunk that reads those words."""
import numpy as np
import pytest

import _oracle as O

TILE, CHUNK = 64, 16


def chunked_scan(mask, n, late_workers):
    """mask: [n, cb] uint64, upper triangle (words left of a row's own block are garbage in the kernel: poisoned here)"""
    cb = (n + TILE - 1) // TILE
    remv_g = [0] * cb
    keep = []
    s_next = [0] * CHUNK
    pending = []   # tickets (rows, word0) not yet applied by the workers
    rows_of = lambda ids, j: np.bitwise_or.reduce(mask[ids, j]) if len(ids) else np.uint64(0)

    def apply(ticket):
        ids, word0 = ticket
        for j in range(word0, cb):
            remv_g[j] |= int(rows_of(ids, j))

    for ci, c0 in enumerate(range(0, cb, CHUNK)):
        nblk = min(CHUNK, cb - c0)
        # tickets issued by chunks <= ci-2 must be complete before this chunk reads its words (the kernel spins on ctl->done);
        # a ticket issued by chunk ci-1 starts at this chunk's END, so it may still be outstanding
        still = []
        for t_ci, ticket in pending:
            if t_ci <= ci - 2 or not late_workers:
                apply(ticket)
            else:
                still.append((t_ci, ticket))
        pending = still
        rm = [remv_g[c0 + w] | s_next[w] for w in range(nblk)]
        s_next = [0] * CHUNK
        full = lambda w: min(n - (c0 + w) * TILE, TILE)
        if all((rm[w] & ((1 << full(w)) - 1)) == (1 << full(w)) - 1 for w in range(nblk)):
            continue
        kept_rows = []
        for bl in range(nblk):
            base, size = (c0 + bl) * TILE, full(bl)
            allm = (1 << size) - 1
            if (rm[bl] & allm) == allm:
                continue
            r, kept = rm[bl] | (~allm & (2 ** 64 - 1)), 0
            for i in range(TILE):
                if not (r >> i) & 1:
                    kept |= 1 << i
                    r |= int(mask[base + i, c0 + bl])
            ids = [base + i for i in range(TILE) if (kept >> i) & 1]
            keep.extend(ids)
            kept_rows.extend(ids)
            for w in range(bl + 1, nblk):
                rm[w] |= int(rows_of(ids, c0 + w))
        word0 = c0 + nblk
        if kept_rows and word0 < cb:
            if word0 + CHUNK < cb:
                pending.append((ci, (kept_rows, word0 + CHUNK)))
            for l in range(min(CHUNK, cb - word0)):
                s_next[l] |= int(rows_of(kept_rows, word0 + l))
    return keep


@pytest.mark.parametrize("n,thresh,seed", [(2500, 0.3, 1), (3000, 0.05, 2), (1100, 0.7, 3), (64 * 16 * 2 + 5, 1e-5, 4)])
@pytest.mark.parametrize("late_workers", [False, True])
def test_chunked_scan_model_equals_greedy_oracle(n, thresh, seed, late_workers):
    boxes = O.synth_boxes(n, 3, seed=seed, rounded=True, extent=64.0)
    want = O.nms(boxes, thresh, 3)
    mask = O.nms_mask(boxes, thresh, 3).astype(np.uint64).reshape(n, -1)
    cb = mask.shape[1]
    # the device mask kernel writes the upper triangle only: poison everything left of a row's own block
    for i in range(n):
        mask[i, : i // TILE] = np.uint64(0xDEADBEEFDEADBEEF)
    got = chunked_scan(mask, n, late_workers)
    assert got == list(map(int, want))
    assert cb == (n + TILE - 1) // TILE
