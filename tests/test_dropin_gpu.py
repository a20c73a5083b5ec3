"""GPU: the drop-in CLIs -- the reference's own host objects linked against our shim +
libsboxgates_b200.so (built by `make -C sboxgates_b200/csrc dropin`):
  oracle/_ref/sboxgates_gpu   node-level shim (lut_search = one device call chain per node)
  oracle/_ref/sboxgates_gpu2  two-function shim (search_5lut / search_7lut)
must write exactly the files the seeded reference writes (tests/golden/xml_names.json: gate count
and Speck fingerprint of the whole graph, state.c:68-125), and every file must pass an independent
functional check (sboxgates_b200/graph.py: tables recomputed from the topology, output gate ==
S-box bit)."""
import glob
import json
import os
import subprocess
import tempfile
import time

import pytest

import _support as S
from sboxgates_b200 import graph as G

pytestmark = pytest.mark.gpu

EXE = os.path.join(S.REF_DIR, "sboxgates_gpu")
EXE2 = os.path.join(S.REF_DIR, "sboxgates_gpu2")
REF = os.path.join(S.REF_DIR, "sboxgates_ref")
needs_exe = pytest.mark.skipif(not (os.path.exists(EXE) and os.path.exists(EXE2)),
                               reason="oracle/_ref/sboxgates_gpu{,2} not built")

# BASELINE.json configs[1] under the two committed seeds.  The reference cannot finish these runs
# (minutes per search_7lut call from n = 20 on), so the names below were produced by THIS repo; what
# backs them: the first 150 search calls of the seed1 run replay bit-exactly against the reference's
# recorded outputs (tests/test_gpu_parity.py, run_rijndael_seed1.bin), both shims and every GPU count
# give the same file, and the file passes the independent functional check below.
RIJNDAEL = {"seed1": "1-031-0000-0-55aa04f1.xml", "seed2": "1-031-0000-0-4a5be130.xml"}


def _run(exe, sbox, cli, seed, tmp, timeout=600, extra_env=None):
    env = dict(os.environ, SBG_SEEDFILE=os.path.join(S.GOLDEN, seed + ".bin"), SBG_SHIM_STATS="1")
    env.update(extra_env or {})
    t0 = time.time()
    res = subprocess.run([exe] + cli + [os.path.join(S.REF_DIR, "sboxes", sbox)], cwd=tmp, env=env,
                         capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, res.stderr[-2000:]
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(tmp, "*.xml"))), \
        time.time() - t0, res.stderr


def _verify(tmp, name, sbox_file, bits):
    sbox, _ = G.load_sbox(os.path.join(S.REF_DIR, "sboxes", sbox_file))
    graph = G.load_graph(os.path.join(tmp, name))
    assert G.verify_graph(graph, sbox, require_bits=bits) == sorted(bits)
    return graph


@needs_exe
@pytest.mark.parametrize("exe", [EXE, EXE2], ids=["node", "two-function"])
def test_dropin_reproduces_reference_graphs(exe):
    names = json.load(open(os.path.join(S.GOLDEN, "xml_names.json")))
    assert len(names) >= 8
    for key, want in sorted(names.items()):
        sbox, *cli, seed = key.split()
        with tempfile.TemporaryDirectory() as tmp:
            got, secs, err = _run(exe, sbox, cli, seed, tmp)
            assert got == want, (key, got, want)
            graph = _verify(tmp, got[-1], sbox, [0])
            assert graph.num_luts == int(got[-1].split("-")[1])
        if sbox == "des_s1.txt":
            # the searches really went through the shim
            assert ("lut_search: " in err) if exe == EXE else ("search_7lut" in err)


@needs_exe
def test_dropin_config1_without_lut_is_the_reference():
    """BASELINE.json configs[0]: `-o 0 des_s1.txt` without --lut never enters the LUT path; the
    drop-in binaries must behave exactly like the reference binary (same seed -> same files), which
    also shows that interposing xorshift1024 (node shim) leaves the host's random stream alone."""
    outs = []
    for exe in (REF, EXE, EXE2):
        with tempfile.TemporaryDirectory() as tmp:
            got, secs, err = _run(exe, "des_s1.txt", ["-o", "0"], "seed1", tmp, timeout=300)
            assert got
            _verify(tmp, got[-1], "des_s1.txt", [0])
        outs.append(got)
    assert outs[0] == outs[1] == outs[2]


@needs_exe
@pytest.mark.parametrize("seed", ["seed1", "seed2"])
def test_dropin_finishes_rijndael_single_output(seed):
    """BASELINE.json configs[1]; the reference does not finish this in an hour.  Both shims must
    write the same file, and it must be a correct circuit for output bit 0 -- checked here from the
    XML alone, on top of the reference's own asserts (sboxgates.h:31-44, lut.c:573-576, 617-621)."""
    files = []
    for exe in (EXE, EXE2):
        with tempfile.TemporaryDirectory() as tmp:
            got, secs, err = _run(exe, "rijndael.txt", ["-l", "-o", "0"], seed, tmp)
            assert len(got) == 1
            graph = _verify(tmp, got[0], "rijndael.txt", [0])
            assert graph.num_inputs == 8 and graph.num_luts == int(got[0].split("-")[1])
            assert secs < 120
        files.append(got[0])
    assert files[0] == files[1] == RIJNDAEL[seed]


@needs_exe
def test_dropin_loads_and_converts_its_own_graphs():
    """--graph / -d / -c on a file the GPU build wrote (state.c:260-411 through the libxml2-free
    reader of the drop-in build): continue a des_s1 graph with a second output, convert it."""
    with tempfile.TemporaryDirectory() as tmp:
        got, _, _ = _run(EXE, "des_s1.txt", ["-l", "-o", "0"], "seed1", tmp)
        first = got[-1]
        got2, _, _ = _run(EXE, "des_s1.txt", ["-l", "-o", "1", "-g", os.path.join(tmp, first)],
                          "seed2", tmp)
        two = [n for n in got2 if n.startswith("2-")]
        assert two, got2
        _verify(tmp, two[-1], "des_s1.txt", [0, 1])
        env = dict(os.environ, SBG_SEEDFILE=os.path.join(S.GOLDEN, "seed1.bin"))
        dot = subprocess.run([EXE, "-d", os.path.join(tmp, two[-1])], cwd=tmp, env=env,
                             capture_output=True, text=True, timeout=60)
        assert dot.returncode == 0 and dot.stdout.lstrip().startswith("digraph")
        cfun = subprocess.run([EXE, "-c", os.path.join(tmp, two[-1])], cwd=tmp, env=env,
                              capture_output=True, text=True, timeout=60)
        assert cfun.returncode == 0 and "lop3" in cfun.stdout.lower()


@needs_exe
def test_dropin_sharded_over_two_devices_gives_the_same_graph():
    """SBG_GPUS=2 with the sharding thresholds at zero: every search is split over two devices
    (one host thread each, lists gathered and merged on the devices); the graph must not change.
    Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    names = json.load(open(os.path.join(S.GOLDEN, "xml_names.json")))
    key = "des_s1.txt -l -o 0 seed1"
    extra = {"SBG_GPUS": "2", "SBG_SHARD_MIN5": "0", "SBG_SHARD_MIN7": "0",
             "SBG_SHARD_MIN_LIST": "0"}
    for exe in (EXE, EXE2):
        with tempfile.TemporaryDirectory() as tmp:
            got, secs, err = _run(exe, "des_s1.txt", ["-l", "-o", "0"], "seed1", tmp, extra_env=extra)
        assert got == names[key]
        assert "sharded search phases" in err
        with tempfile.TemporaryDirectory() as tmp:
            got, secs, err = _run(exe, "rijndael.txt", ["-l", "-o", "0"], "seed1", tmp,
                                  extra_env=extra)
        assert got == [RIJNDAEL["seed1"]]
