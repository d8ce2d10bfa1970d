"""Multi-rank parity of the frame-sharded pipeline as a driver-run test: N ranks of StreamingExtractor must give
the 1-rank answer -- per-frame result rows, unique count, tempo_count identical -- on dense glyph masks, so that every hand-off
carries thousands of uniques and their crops.

The peer-memory ring (csrc/p2p.cu: CUDA-IPC mailboxes + stream memory operations) is exercised on ONE GPU by two processes that
share cuda:0; the NCCL hand-off needs two devices (NCCL refuses two ranks on one) and is skipped on a 1-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from tests.conftest import REPO

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_ring(world, extra, timeout=600):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   AM_B200_HANDOFF_TIMEOUT="60")
        procs.append(subprocess.Popen([sys.executable, os.path.join(REPO, "tools", "ring_check.py")] + extra, env=env, cwd=REPO,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=timeout)[0])
    finally:
        for p in procs:                                     # the exact processes this test started
            if p.poll() is None:
                p.kill()
    return [p.returncode for p in procs], outs


@pytest.mark.parametrize("masks", ["glyph", "fcn"])
def test_two_rank_p2p_ring_on_one_gpu_equals_one_rank(masks):
    codes, outs = _run_ring(2, ["--same-device", "--handoff", "p2p", "--masks", masks, "--hw", "720x1280", "--batch", "4", "--rounds", "4"])
    assert codes == [0, 0], "\n".join(outs)
    line = [l for l in outs[0].splitlines() if l.startswith("ring_check")]
    assert line and "IDENTICAL" in line[0], outs[0]
    if masks == "glyph":                                    # the hand-off really carried a dense active set
        uniques = int(line[0].split("uniques=")[1].split()[0])
        assert uniques > 2000, line[0]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="the NCCL hand-off needs two devices")
@pytest.mark.parametrize("handoff", ["p2p", "nccl"])
def test_two_gpu_ring_equals_one_rank(handoff):
    codes, outs = _run_ring(2, ["--handoff", handoff, "--masks", "glyph", "--hw", "720x1280", "--batch", "4", "--rounds", "4"])
    assert codes == [0, 0], "\n".join(outs)
    assert any(l.startswith("ring_check") and "IDENTICAL" in l for l in outs[0].splitlines()), outs[0]
