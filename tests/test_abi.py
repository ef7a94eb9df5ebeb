"""CPU tests of the C-ABI surface: the library loads, exports every symbol include/burg_b200.h declares,
and fails loudly (never falls back) without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from burg_toolkit_b200 import _native as nat


def _declared_symbols():
    header = open(os.path.join(ROOT, 'include', 'burg_b200.h')).read()
    return sorted(set(re.findall(r'^BT_API\s+[\w\s\*]+?\b(bt_\w+)\(', header, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in burg_b200.h but not exported'
    assert sorted(nat.EXPORTED_SYMBOLS) == declared, 'ctypes signature table out of sync with the header'
    assert lib.bt_version() == 100


def test_struct_layouts_match_header():
    assert nat.HIT_DTYPE.itemsize == 72
    assert ctypes.sizeof(nat.AgsParams) == 48
    assert ctypes.sizeof(nat.GripperPart) == 56
    assert ctypes.sizeof(nat.MeshInfo) == 8 * 4 + 8 * 2 + 4 * 6 + 4 * 2


def test_argument_validation_needs_no_device():
    lib = nat.load()
    out = ctypes.c_void_p()
    rc = lib.bt_mesh_create(None, 0, None, 0, None, 0, ctypes.byref(out))
    assert rc == -1 and b'null pointer' in lib.bt_last_error()


def test_tree_build_selector_validates_its_argument():
    """bt_set_bvh_build is process-wide state and needs no device: 0 / 1 / 2 are accepted, anything else is refused"""
    from burg_toolkit_b200 import device
    lib = nat.load()
    try:
        for kind in (0, 1, 2):
            assert lib.bt_set_bvh_build(kind) == 0
        assert lib.bt_set_bvh_build(3) == -1 and b'kind must be' in lib.bt_last_error()
        assert lib.bt_set_bvh_build(-1) == -1
        with pytest.raises(KeyError):
            device.set_bvh_build('sah')
        kind = ctypes.c_int32()
        assert lib.bt_mesh_bvh_kind(None, ctypes.byref(kind), None, None) == -1 and b'null pointer' in lib.bt_last_error()
    finally:
        device.set_bvh_build('ploc')


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from burg_toolkit_b200 import metrics, sampling, mesh
    from burg_toolkit_b200.grasp import GraspSet
    gs = GraspSet(n=4)
    with pytest.raises(nat.NativeError):
        metrics.coverage(gs, gs)
    ags = sampling.AntipodalGraspSampler()
    ags.mesh, ags.verbose = mesh.create_icosphere(1, 0.03), False
    with pytest.raises(nat.NativeError):
        ags.sample(10)
    v = np.zeros((3, 3)); v[1, 0] = v[2, 1] = 1
    f = np.array([[0, 1, 2]], dtype=np.int32)
    out = ctypes.c_void_p()
    rc = nat.load().bt_mesh_create(v.ctypes.data, 3, f.ctypes.data, 1, None, 0, ctypes.byref(out))
    assert rc in (-4, -2)          # BT_E_NODEVICE / BT_E_CUDA -- never a silent CPU path


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'burg_toolkit_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
                assert 'oracle/' not in src.replace('oracle/port.py', '') or fn.endswith('.md'), fn


def test_ctypes_mirrors_agree_with_the_c_header(tmp_path):
    """sizeof / offsetof of every struct that crosses the C ABI, as gcc sees include/burg_b200.h, against the ctypes mirror"""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    structs = {'bt_ags_params': nat.AgsParams, 'bt_mesh_info': nat.MeshInfo, 'bt_gripper_part': nat.GripperPart,
               'bt_pipeline_config': nat.PipelineConfig, 'bt_pipeline_buffers': nat.PipelineBuffers}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "burg_b200.h")}"', 'int main(void) {']
    for cname, mirror in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in mirror._fields_:
            lines.append(f'  printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    lines.append('  printf("bt_hit %zu\\n", sizeof(bt_hit));  return 0; }')
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-o', str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, mirror in structs.items():
        assert int(out[cname]) == ctypes.sizeof(mirror), cname
        for field, _ in mirror._fields_:
            assert int(out[f'{cname}.{field}']) == getattr(mirror, field).offset, f'{cname}.{field}'
    assert int(out['bt_hit']) == nat.HIT_DTYPE.itemsize
