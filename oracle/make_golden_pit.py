"""Golden vectors for a9 (pit_loss): the reference's OWN cell 28 of uPIT_baseline.ipynb, executed as it stands.

TEST INFRASTRUCTURE.  TensorFlow is not installed in this image, so the cell's source text (read from
/root/reference at generation time, never copied into the repository) runs against `TfShim`: the ten `tf.*`
functions the cell calls, each restated on torch with TensorFlow's documented semantics (slice with -1 sizes,
sequence_mask, tile, reduce_sum over one axis, broadcasting).  That pins `oracle.signal_path.pit_mse` to the
reference's own slicing, masking, reduction order and permutation rule instead of to my reading of them; the
gradient comes from torch autograd through the very same executed code (TF's autodiff would give the same
derivative of the same expression).  What stays unpinned is only the shim's ten one-liners.

usage (authoring container):  python oracle/make_golden_pit.py   ->  tests/golden/pit_golden.npz
"""
from __future__ import annotations

import io
import json
import os
from contextlib import redirect_stdout

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NOTEBOOK = "/root/reference/uPIT_baseline.ipynb"
OUT = os.path.join(os.path.dirname(_HERE), "tests", "golden", "pit_golden.npz")


class TfShim:
    """The subset of the TensorFlow API that cell 28 uses, on torch tensors."""

    def __init__(self):
        import torch

        self.t = torch
        self.float32 = torch.float32
        self.float64 = torch.float64

    def shape(self, x):
        return tuple(x.shape)

    def slice(self, x, begin, size):                       # size -1: everything from `begin` to the end of the axis
        idx = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
        return x[idx]

    def squeeze(self, x):
        return x.squeeze()

    def sequence_mask(self, lengths, maxlen):
        # array_ops.sequence_mask: row = range(maxlen) in maxlen's integer dtype, `lengths` CAST to that dtype
        # (a float length truncates toward zero), mask[b, t] = row[t] < lengths[b]
        steps = self.t.arange(int(maxlen), dtype=self.t.int64)
        return steps < lengths.to(self.t.int64).unsqueeze(-1)

    def cast(self, x, dtype):
        return x.to(x.dtype if (dtype == self.float32 and x.dtype == self.float64) else dtype)   # golden runs stay float64

    def expand_dims(self, x, axis):
        return x.unsqueeze(axis)

    def tile(self, x, multiples):
        return x.repeat(*[int(m) for m in multiples])

    def pow(self, x, p):
        return x ** p

    def reduce_sum(self, x, axis=None):
        return x.sum() if axis is None else x.sum(dim=axis)

    def reduce_mean(self, x, axis=None):
        return x.mean() if axis is None else x.mean(dim=axis)


def reference_pit_source():
    """The live (uncommented) definition of pit_with_outputsize in the notebook."""
    nb = json.load(open(NOTEBOOK))
    for cell in nb["cells"]:
        src = "".join(cell["source"])
        if "def pit_with_outputsize" in src and "sequence_mask" in src:
            return src
    raise RuntimeError("cell 28 not found")


def load_reference_pit():
    ns = {"tf": TfShim()}
    exec(compile(reference_pit_source(), NOTEBOOK + ":cell28", "exec"), ns)
    return ns["pit_with_outputsize"]


def run_reference(y_true, y_pred, output_size):
    """(loss, d loss / d y_pred) of the reference's cell, float64."""
    import torch

    pit = load_reference_pit()(output_size)
    yt = torch.from_numpy(np.asarray(y_true, dtype=np.float64))
    yp = torch.from_numpy(np.asarray(y_pred, dtype=np.float64)).requires_grad_(True)
    with redirect_stdout(io.StringIO()):                    # the cell prints tf.shape(mask)
        loss = pit(yt, yp)
    loss.backward()
    return float(loss.detach()), yp.grad.numpy().copy()


def cases():
    """Seeded batches: ragged lengths (incl. a fractional one: int() truncation), a tie, swapped speakers."""
    rng = np.random.default_rng(28)
    out = {}
    for name, (n_batch, n_time, feat) in {"small": (3, 7, 5), "cfg": (4, 40, 129)}.items():
        labels = rng.random((n_batch, n_time, 2 * feat))
        pred = rng.random((n_batch, n_time, 2 * feat))
        lengths = rng.integers(max(1, n_time // 2), n_time + 1, size=n_batch).astype(np.float64)
        if name == "small":
            lengths[0] = n_time
            lengths[1] = 4.7                                  # sequence_mask casts to int: 4 frames
            pred[2] = np.concatenate([labels[2][:, feat:], labels[2][:, :feat]], axis=1) + 0.01 * rng.random((n_time, 2 * feat))
        y_true = np.concatenate([labels, np.zeros((n_batch, 1, 2 * feat))], axis=1)
        y_true[:, n_time, 0] = lengths
        out[name] = (y_true, pred, feat)
    # exact tie: both speakers' predictions and labels identical -> cost1 == cost2 -> idx = 0
    lab = rng.random((2, 6, 4))
    lab[:, :, 2:] = lab[:, :, :2]
    pr = rng.random((2, 6, 4))
    pr[:, :, 2:] = pr[:, :, :2]
    yt = np.concatenate([lab, np.zeros((2, 1, 4))], axis=1)
    yt[:, 6, 0] = (6, 3)
    out["tie"] = (yt, pr, 2)
    return out


def main():
    blob = {}
    for name, (y_true, y_pred, feat) in cases().items():
        loss, grad = run_reference(y_true, y_pred, feat)
        blob[name + "_y_true"] = y_true
        blob[name + "_y_pred"] = y_pred
        blob[name + "_feat"] = np.int64(feat)
        blob[name + "_loss"] = np.float64(loss)
        blob[name + "_grad"] = grad
        print("%-6s loss %.12f  |grad| %.6f" % (name, loss, np.abs(grad).sum()))
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
