"""The C-ABI library loads without a GPU, exports every symbol include/seekr2_b200.h declares, and the
ctypes mirrors of its structs have the layout gcc gives the header."""
import ctypes as C
import os
import re
import subprocess

import pytest

import helpers as H
from helpers import L

HEADER = open(L.HEADER_PATH).read()


def test_library_exports_every_declared_symbol():
    declared = set(re.findall(r"^(?:int|const char\*)\s+(seekr2_b200_\w+)\s*\(", HEADER, flags=re.M))
    assert len(declared) >= 20
    assert declared == set(L.SYMBOLS), (declared ^ set(L.SYMBOLS))
    lib = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.seekr2_b200_abi_version() == L.ABI_VERSION


def test_struct_layout_matches_gcc(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "seekr2_b200.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu\\n", sizeof(seekr2_b200_boundary), sizeof(seekr2_b200_config),'
                   ' sizeof(seekr2_b200_buffers), sizeof(seekr2_b200_record), sizeof(seekr2_b200_anchor_state));\n'
                   'printf("%zu %zu %zu %zu\\n", offsetof(seekr2_b200_config, group_offsets), offsetof(seekr2_b200_config, boundaries),'
                   ' offsetof(seekr2_b200_config, ring_capacity), offsetof(seekr2_b200_anchor_state, elber_values));\n'
                   'printf("%zu %zu\\n", offsetof(seekr2_b200_boundary, bitcode), offsetof(seekr2_b200_record, time));\nreturn 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.dirname(L.HEADER_PATH), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    want = [C.sizeof(L.Boundary), C.sizeof(L.Config), C.sizeof(L.Buffers), C.sizeof(L.Record), C.sizeof(L.AnchorState),
            L.Config.group_offsets.offset, L.Config.boundaries.offset, L.Config.ring_capacity.offset,
            L.AnchorState.elber_values.offset, L.Boundary.bitcode.offset, L.Record.time.offset]
    assert got == want


def test_header_is_plain_c_with_reference_citations():
    assert 'extern "C"' in HEADER and "#include <torch" not in HEADER and "at::" not in HEADER and "std::" not in HEADER
    for cite in ("Seekr2Kernels.h:53-110", "CudaSeekr2Kernels.cpp:82-366", "langevin.cu:7-179"):
        assert cite in HEADER


def test_config_validation_needs_no_gpu():
    """create() rejects bad configurations before touching CUDA; error text is retrievable."""
    cfg = L.Config()
    h = C.c_void_p()
    cfg.abi_version = 99
    assert L.lib.seekr2_b200_create(C.byref(cfg), C.byref(h)) == L.ERR_INVALID
    assert b"ABI version" in L.lib.seekr2_b200_last_error()
    cfg.abi_version = L.ABI_VERSION
    cfg.kind, cfg.precision, cfg.num_anchors, cfg.num_atoms, cfg.padded_num_atoms = L.KIND_MMVT, 7, 1, 1, 32
    assert L.lib.seekr2_b200_create(C.byref(cfg), C.byref(h)) == L.ERR_INVALID
    assert b"precision" in L.lib.seekr2_b200_last_error()
    with pytest.raises(L.Seekr2Error):
        L.check(L.ERR_INVALID)


def test_elber_missing_force_group_is_rejected():
    """CudaSeekr2Kernels.cpp:415-436: every Elber milestone needs a Force in its force group."""
    syn = H.toy_system()
    integ = H.s2.ElberLangevinIntegrator(300.0, 1.0, 0.002, "")
    integ.addSrcMilestoneGroup(1)
    integ.addDestMilestoneGroup(9)          # no boundary carries force group 9
    cfg = integ.makeConfig(syn.system, L.MIXED)
    h = C.c_void_p()
    assert L.lib.seekr2_b200_create(C.byref(cfg), C.byref(h)) == L.ERR_NO_FORCE_GROUP
    assert b"no force groups used to detect Elber boundary crossings" in L.lib.seekr2_b200_last_error()


def test_float_boundary_flag_is_validated_before_cuda():
    """FLAG_CV_FLOAT restates OpenMM's single-precision evaluation: meaningless in double precision, and defined for spheres and
    planes only -- rejected by create() before any device work."""
    syn = H.toy_system()
    integ = H.make_mmvt_integrator(syn, "")
    integ.setBoundaryValueInFloat(True)
    h = C.c_void_p()
    cfg = integ.makeConfig(syn.system, L.DOUBLE)
    assert L.lib.seekr2_b200_create(C.byref(cfg), C.byref(h)) == L.ERR_INVALID
    assert b"FLAG_CV_FLOAT" in L.lib.seekr2_b200_last_error()


def test_strict_boundary_parameters_in_the_harness():
    """A name in the boundary expression that is not a per-bond parameter is an error, not a silent default."""
    syn = H.toy_system()
    f = syn.boundary_force
    f.param_names[f.param_names.index("radius")] = "rad"
    integ = H.make_mmvt_integrator(syn, "")
    with pytest.raises(L.Seekr2Error) as err:
        integ.makeConfig(syn.system, L.MIXED)
    assert "not a per-bond parameter" in str(err.value)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    syn = H.toy_system()
    with pytest.raises(L.Seekr2Error):
        H.s2.Context(syn.system, H.make_mmvt_integrator(syn, ""), {"CudaPrecision": "mixed"})


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: importing the package, building a workload and lowering a
    configuration must not load it (bench.py's GPU arm relies on this)."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, importlib.abc
        class Guard(importlib.abc.MetaPathFinder):
            def find_spec(self, name, path, target=None):
                if name == 'oracle' or name.startswith('oracle.'):
                    raise ImportError('oracle imported by the product path')
        sys.meta_path.insert(0, Guard())
        sys.path.insert(0, %r)
        import seekr2_openmm_plugin_b200 as s2
        from seekr2_openmm_plugin_b200 import synthetic, multi, _lib
        import bench
        syn, pos, f, sh = bench.build_workload(2, 300, 1)
        integ = synthetic.make_mmvt_integrator(syn, '')
        cfg = integ.makeConfig(syn.system, _lib.MIXED, 2)
        assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)
        print('clean')
    """) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "clean" in out.stdout, out.stderr[-2000:]
