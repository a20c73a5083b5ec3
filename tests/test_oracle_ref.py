"""CPU, only where oracle/_ref/ was built (the build container, and the GPU box through the
snapshot): the oracle against the reference's own object code LIVE, and the reference CLI against the
committed golden file names.  Skipped when oracle/_ref/ is absent."""
import glob
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import _support as S

pytestmark = pytest.mark.skipif(not S.ref_available(), reason="oracle/_ref not built")


def test_struct_layout_the_shim_assumes():
    import ctypes as C
    lib = S.ref_lib()
    a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.sbgref_sizes(C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    assert (a.value, b.value, c.value, d.value) == (32, 64, 32032, 32)   # state.h:64-88
    # the node-level shim (lut_search) also reads three fields of `options` (sboxgates.h:49-66)
    e = C.c_int()
    lib.sbgref_options_layout(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e))
    assert (a.value, b.value, c.value, d.value, e.value) == (2019, 2018, 9756, 9760, 24)


def test_oracle_equals_reference_on_random_cases():
    sbox = S.rijndael_sbox()
    rs = np.random.RandomState(123)
    for i in range(30):
        n = int(rs.choice([7, 8, 9, 10, 11]))
        tabs = S.synthetic_state(n, seed=3000 + i, num_inputs=min(8, n))
        pos = rs.choice(256, int(rs.choice([6, 10, 16, 24, 40])), replace=False)
        mask = np.zeros(4, dtype=np.uint64)
        for p in pos:
            mask[p >> 6] |= np.uint64(1) << np.uint64(p & 63)
        tgt = S.sbox_target(sbox, int(rs.randint(0, 8)))
        inb = [int(rs.randint(0, min(8, n)))] if i % 3 == 0 else []
        for which in (5, 7):
            seed = rs.bytes(128)
            a = S.OrcRng.from_seed(seed)
            b = S.OrcRng.from_seed(seed)
            found_o, ret_o, _ = S.oracle_search(which, tabs, tgt, mask, inb, a)
            found_r, ret_r, draws = S.ref_search(which, tabs, tgt, mask, inb, b)
            assert (found_o, ret_o) == (found_r, ret_r), (i, which)
            assert a.draws == draws and a.words() == b.words() and a.p == b.p


def test_reference_cli_reproduces_golden_file_names():
    names = json.load(open(os.path.join(S.GOLDEN, "xml_names.json")))
    for key in ("crypto1_fa.txt -l seed1", "crypto1_fc.txt -l seed2"):
        sbox, *cli, seed = key.split()
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, SBG_SEEDFILE=os.path.join(S.GOLDEN, seed + ".bin"))
            subprocess.run([os.path.join(S.REF_DIR, "sboxgates_ref")] + cli
                           + [os.path.join(S.REF_DIR, "sboxes", sbox)], cwd=tmp, env=env, check=True,
                           stdout=subprocess.DEVNULL)
            got = sorted(os.path.basename(p) for p in glob.glob(os.path.join(tmp, "*.xml")))
        assert got == names[key]


def test_rijndael_table_is_the_aes_sbox():
    txt = open(os.path.join(S.REF_DIR, "sboxes", "rijndael.txt")).read().split()
    assert [int(x, 16) for x in txt] == S.rijndael_sbox()


def test_saved_graph_loads_back_and_converts():
    """The XML the reference writes is read back by its own loader through our mini XML reader
    (sboxgates_b200/csrc/xmlmini) and converts to DOT -- exercises --graph/-d in the oracle build."""
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, SBG_SEEDFILE=os.path.join(S.GOLDEN, "seed1.bin"))
        exe = os.path.join(S.REF_DIR, "sboxgates_ref")
        subprocess.run([exe, "-l", os.path.join(S.REF_DIR, "sboxes", "crypto1_fc.txt")], cwd=tmp,
                       env=env, check=True, stdout=subprocess.DEVNULL)
        xml = glob.glob(os.path.join(tmp, "*.xml"))[0]
        dot = subprocess.run([exe, "-d", xml], cwd=tmp, env=env, check=True, capture_output=True,
                             text=True).stdout
        assert "digraph sbox" in dot and "-> gt" in dot
        # the Python loader (sboxgates_b200/graph.py) reads the reference's own file and confirms
        # the circuit, and a 2-input-gate graph (no --lut) as well
        from sboxgates_b200 import graph as G
        sbox, _ = G.load_sbox(os.path.join(S.REF_DIR, "sboxes", "crypto1_fc.txt"))
        g = G.load_graph(xml)
        assert G.verify_graph(g, sbox, require_bits=[0]) == [0]
        assert g.num_luts == int(os.path.basename(xml).split("-")[1])
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([exe, "-o", "0", os.path.join(S.REF_DIR, "sboxes", "des_s1.txt")], cwd=tmp,
                       env=env, check=True, stdout=subprocess.DEVNULL, timeout=300)
        xml = sorted(glob.glob(os.path.join(tmp, "*.xml")))[-1]
        sbox, _ = G.load_sbox(os.path.join(S.REF_DIR, "sboxes", "des_s1.txt"))
        assert G.verify_graph(G.load_graph(xml), sbox, require_bits=[0]) == [0]
