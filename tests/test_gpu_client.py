"""The reference's PqaClient learner loop (ProbQA/PqaClient/PqaClient.cpp:69-245) re-expressed in C++ over the C ABI
(clients/pqa_client.cpp, reference symbols only), run against the B200 engine: it must complete, write progress lines in
the reference's format, learn (top-1 precision rises), and its questions/s column -- the only throughput the reference
publishes (mean 301 questions/s) -- is reported."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu


def test_pqa_client_learner_loop(tmp_path):
    from probqa_b200 import build
    exe = build.build_client()
    progress = str(tmp_path / "progress.txt")
    out = subprocess.run([exe, "--trainings", "6000", "--learners", "48", "--report-every", "1024", "--progress", progress,
                          "--kb-dir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["failed"] is False and line["questions_asked"] > 6000
    rows = [r.split("\t") for r in open(progress).read().strip().splitlines()]
    assert len(rows) == 5 and all(len(r) == 6 for r in rows)              # trainings 1024 .. 5120
    assert [int(r[0]) for r in rows] == [1024 * (x + 1) for x in range(5)]
    precision = [float(r[2]) for r in rows]
    assert precision[-1] > precision[0]                                    # it learns
    assert float(rows[-1][5]) > 301.0                                      # reference-published questions/s (unrecorded CPU)
    assert os.path.exists(str(tmp_path / "dichotomy001024.kb"))           # KB snapshots like the reference's (PqaClient.cpp:92-98)
    print("pqa_client:", json.dumps(line))
