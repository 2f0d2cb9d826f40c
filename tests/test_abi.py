"""CPU: the C-ABI library loads and exports every symbol include/fme_b200.h declares; record layouts match;
the product refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from common import fme

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fme_b200.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    inline = set(re.findall(r"static\s+inline\s+[a-z0-9_ ]+?\b(fme_[a-z0-9_]+)\s*\(", text))   # header-only helpers
    return sorted(set(re.findall(r"\b(fme_[a-z0-9_]+)\s*\(", text)) - inline)


def test_header_symbols_are_exported():
    import __graft_entry__ as ge
    ge.build()
    lib = fme.load_library()
    names = header_functions()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(fme.EXPORTS) == names


def test_record_layouts_match_header():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "fme_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(fme_pu), offsetof(fme_pu, mvIntX), offsetof(fme_pu, err),
         sizeof(fme_result), offsetof(fme_result, cost), offsetof(fme_result, nnClass), sizeof(fme_mc_pu), sizeof(fme_config));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        vals = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    dt, rt = fme.PU_DTYPE, fme.RESULT_DTYPE
    assert vals == [dt.itemsize, dt.fields["mvIntX"][1], dt.fields["err"][1], rt.itemsize, rt.fields["cost"][1],
                    rt.fields["nnClass"][1], fme.MC_PU_DTYPE.itemsize, C.sizeof(fme.FmeConfig)]


def test_oracle_record_layout_matches_product():
    import oracle_bindings as ob
    assert ob.PU_DTYPE == fme.PU_DTYPE and ob.RESULT_DTYPE == fme.RESULT_DTYPE


def test_no_cpu_fallback():
    """Without a CUDA device the engine must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fme.FmeError) as e:
        fme.Fme(64, 64)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "hm16.9-nn_fme_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "libfme_oracle" not in text and "libhmref" not in text and "oracle_bindings" not in text, f


def test_hm_adaptor_compiles_against_reference_headers():
    """The header-only C++ adaptor (reference signatures over the C ABI) must compile inside the reference tree."""
    ref = "/root/reference/source/Lib"
    if not os.path.isdir(ref):
        pytest.skip("reference headers not available on this box")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write('#include "fme_hm_adaptor.h"\nint main() { FmeHmAdaptor a; (void)a; return 0; }\n')
        subprocess.check_call(["g++", "-std=gnu++11", "-fsyntax-only", "-w", "-I", ref, "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "hm16.9-nn_fme_b200", "adaptor"), src])


def test_compact_record_pack_helper_matches_python_mirror():
    """fme_pu_compact_pack (header-only, what an adaptor calls per PU) == pu_list.compact_of: 44-byte layout, 24-bit
    little-endian grid values, return value 1 exactly for the records that need the fme_err_grid list."""
    rng = np.random.default_rng(11)
    recs = np.zeros(64, fme.PU_DTYPE)
    for f in ("x", "y", "mvIntX", "mvIntY", "mvPredX", "mvPredY"):
        recs[f] = rng.integers(-2000, 2000, len(recs))
    recs["w"], recs["h"], recs["refSlot"], recs["flags"] = 16, 8, rng.integers(0, 4, len(recs)), rng.integers(0, 2, len(recs))
    recs["err"] = rng.integers(0, 1 << 24, recs["err"].shape).astype(np.uint32)
    recs["err"][::7, 3] = rng.integers(1 << 24, 1 << 32, len(recs[::7]), dtype=np.uint64).astype(np.uint32)
    src = r'''
#include <stdio.h>
#include "fme_b200.h"
int main(void) {
  fme_pu p; fme_pu_compact c;
  if (sizeof(fme_pu_compact) != 44) return 2;
  while (fread(&p, sizeof p, 1, stdin) == 1) {
    unsigned char big = (unsigned char)fme_pu_compact_pack(&p, &c);
    fwrite(&c, sizeof c, 1, stdout);
    fwrite(&big, 1, 1, stdout);
  }
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.run([os.path.join(d, "t")], input=recs.tobytes(), stdout=subprocess.PIPE, check=True).stdout
    got = np.frombuffer(out, np.uint8).reshape(len(recs), 45)
    comp, big = fme.pu_list.compact_of(recs)
    assert np.array_equal(got[:, :44], comp.view(np.uint8).reshape(len(recs), 44))
    flag = np.zeros(len(recs), np.uint8); flag[big["pu"]] = 1
    assert np.array_equal(got[:, 44], flag) and len(big) == len(recs[::7])
    assert np.array_equal(big["err"], recs["err"][big["pu"]])
