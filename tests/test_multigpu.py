"""On-hardware multi-GPU correctness (SURVEY.md 4 item v): the same utterance set at 1 and at 2 (4, 8) ranks
over NCCL must give identical per-utterance results and sums.  Needs >= 2 GPUs; skipped on a 1-GPU box
(the host logic is covered by the gloo tests in test_distributed_cpu.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "nccl_worker.py"), path]
    subprocess.run(cmd, check=True, timeout=600)
    return np.load(path)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_results_identical_to_one_rank(world, tmp_path):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    one = _run(1, str(tmp_path / "w1.npz"))
    many = _run(world, str(tmp_path / ("w%d.npz" % world)))
    # permutations are exact; per-utterance values may differ in the last bits only because the strip
    # partition (which frames share a transform) depends on the batch shape of the call
    assert np.array_equal(one["pit_perm"], many["pit_perm"])
    for key in ("pit_loss", "si_best", "sdr_best"):
        assert one[key].shape == many[key].shape == (22,)
        assert np.allclose(one[key], many[key], rtol=1e-5, atol=1e-4), key
    # all-reduced sums == the sum of the gathered per-utterance values (fixed utterance order), to 1e-12
    for i, key in enumerate(("pit_loss", "si_best", "sdr_best")):
        assert abs(many["sums"][i] - many[key].sum()) <= 1e-12 * abs(many[key].sum())
        assert abs(many["sums"][i] - one["sums"][i]) <= 1e-5 * abs(one["sums"][i])
    assert many["sums"][3] == one["sums"][3] == 22
    # the one-sided push (sep_fused_separate_push_f32) delivers the same sums as the NCCL all-reduce
    ranks_with_work = min(world, 22)
    assert int(many["push_arrived"]) == ranks_with_work and int(one["push_arrived"]) == 1
    assert np.allclose(many["push_sums"], many["sums"], rtol=1e-12, atol=0) and many["push_untouched"] == 0.0
    assert np.array_equal(one["push_sums"], one["sums"])


def test_scoring_sharded_bit_identical(tmp_path):
    """The ragged scorer works per utterance (no batch-shape dependence): sharded over 2 ranks, the gathered
    per-utterance SI-SDR values are BIT-identical to the 1-rank run and the reduced sums agree to 1e-12."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    code = os.path.join(HERE, "nccl_score_worker.py")
    outs = []
    for world in (1, 2):
        path = str(tmp_path / ("s%d.npz" % world))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), code, path]
        subprocess.run(cmd, check=True, timeout=600)
        outs.append(np.load(path))
    assert np.array_equal(outs[0]["si_best"], outs[1]["si_best"])
    assert np.array_equal(outs[0]["si_perm"], outs[1]["si_perm"])
    assert np.array_equal(outs[0]["sdr_best"], outs[1]["sdr_best"])
    for i in range(2):
        assert abs(outs[0]["sums"][i] - outs[1]["sums"][i]) <= 1e-12 * abs(outs[0]["sums"][i])
    assert outs[0]["sums"][2] == outs[1]["sums"][2] == 40
