"""GPU: the drop-in CLI -- the reference's own host objects linked against lut_shim.c +
libsboxgates_b200.so (oracle/_ref/sboxgates_gpu, built by `make -C sboxgates_b200/csrc dropin`) --
must write exactly the files the seeded reference writes (tests/golden/xml_names.json: gate count
and Speck fingerprint of the whole graph, state.c:68-125)."""
import glob
import json
import os
import subprocess
import tempfile
import time

import pytest

import _support as S

pytestmark = pytest.mark.gpu

EXE = os.path.join(S.REF_DIR, "sboxgates_gpu")


def _run(sbox, cli, seed, tmp, timeout=600):
    env = dict(os.environ, SBG_SEEDFILE=os.path.join(S.GOLDEN, seed + ".bin"), SBG_SHIM_STATS="1")
    t0 = time.time()
    res = subprocess.run([EXE] + cli + [os.path.join(S.REF_DIR, "sboxes", sbox)], cwd=tmp, env=env,
                         capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, res.stderr[-2000:]
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(tmp, "*.xml"))), \
        time.time() - t0, res.stderr


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/sboxgates_gpu not built")
def test_dropin_reproduces_reference_graphs():
    names = json.load(open(os.path.join(S.GOLDEN, "xml_names.json")))
    assert len(names) >= 8
    for key, want in sorted(names.items()):
        sbox, *cli, seed = key.split()
        with tempfile.TemporaryDirectory() as tmp:
            got, secs, err = _run(sbox, cli, seed, tmp)
        assert got == want, (key, got, want)
        if sbox == "des_s1.txt":
            assert "search_7lut" in err   # the searches really went through the shim


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/sboxgates_gpu not built")
def test_dropin_finishes_rijndael_single_output():
    """BASELINE.json configs[1]; the reference does not finish this in an hour.  The graph must be
    a correct circuit for output bit 0: the reference's own asserts re-verify every returned gate
    (sboxgates.h:31-44, lut.c:573-576, 617-621) and would abort otherwise."""
    with tempfile.TemporaryDirectory() as tmp:
        got, secs, err = _run("rijndael.txt", ["-l", "-o", "0"], "seed1", tmp)
    assert len(got) == 1 and got[0].startswith("1-0")
    assert got[0] == "1-031-0000-0-55aa04f1.xml"   # stable across every kernel rewrite of round 1
    assert secs < 120


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/sboxgates_gpu not built")
def test_dropin_sharded_over_two_devices_gives_the_same_graph():
    """SBG_GPUS=2 with the sharding thresholds at zero: every search is split over two devices
    (one host thread each); the graph must not change.  Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    names = json.load(open(os.path.join(S.GOLDEN, "xml_names.json")))
    key = "des_s1.txt -l -o 0 seed1"
    env_extra = {"SBG_GPUS": "2", "SBG_SHARD_MIN5": "0", "SBG_SHARD_MIN7": "0",
                 "SBG_SHARD_MIN_LIST": "0"}
    old = {k: os.environ.get(k) for k in env_extra}
    os.environ.update(env_extra)
    try:
        with tempfile.TemporaryDirectory() as tmp:
            got, secs, err = _run("des_s1.txt", ["-l", "-o", "0"], "seed1", tmp)
        assert got == names[key]
        assert "sharded search phases" in err
        with tempfile.TemporaryDirectory() as tmp:
            got, secs, err = _run("rijndael.txt", ["-l", "-o", "0"], "seed1", tmp)
        assert got == ["1-031-0000-0-55aa04f1.xml"]
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
