"""
CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/tnmf_b200.h
declares, and its argument validation / shape logic (no kernel launches) follows the reference's conventions.
"""
import ctypes
import os
import re

import pytest

from tnmf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    _lib.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    header = open(os.path.join(ROOT, 'include', 'tnmf_b200.h')).read()
    declared = set(re.findall(r'\b(tnmf_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.tnmf_abi_version() == int(re.search(r'#define TNMF_ABI_VERSION (\d+)', header).group(1))


def test_struct_layout_matches_header():
    # 8 int32 + 2*3 int32 + 2 int64 + 2 int32 = 80 bytes, no padding surprises
    assert ctypes.sizeof(_lib.Problem) == 8 * 4 + 6 * 4 + 2 * 8 + 2 * 4
    assert _lib.Problem.h_stride_n.offset == 56 and _lib.Problem.flags.offset == 72


@pytest.mark.parametrize('mode,expected', [('valid', (24, 18)), ('full', (16, 16)), ('circular', (20, 17))])
def test_transform_shape_follows_reference(lib, mode, expected):
    """tnmf/backends/_Backend.py:60-73."""
    p = _lib.make_problem(3, 2, 4, (20, 17), (5, 2), _lib.TNMF_F64, mode)
    t = (ctypes.c_int32 * 3)()
    assert lib.tnmf_transform_shape(ctypes.byref(p), t) == 0
    assert (t[0], t[1]) == expected


def test_argument_validation(lib):
    p = _lib.make_problem(3, 2, 4, (20,), (5,), _lib.TNMF_F32)
    assert lib.tnmf_reconstruct(ctypes.byref(p), None, None, None, None, 0, None) == _lib.TNMF_EINVAL
    assert lib.tnmf_workspace_bytes(ctypes.byref(p)) % 256 == 0 and lib.tnmf_workspace_bytes(ctypes.byref(p)) > 0
    p.mode = 7
    assert lib.tnmf_workspace_bytes(ctypes.byref(p)) == 0
    bad = _lib.make_problem(3, 2, 4, (4,), (5,), _lib.TNMF_F32, 'full')           # atom larger than sample
    t = (ctypes.c_int32 * 3)()
    assert lib.tnmf_transform_shape(ctypes.byref(bad), t) == _lib.TNMF_EINVAL
    with pytest.raises(ValueError):
        _lib.make_problem(1, 1, 1, (4,), (2,), 0, 'reflect')
    with pytest.raises(NotImplementedError):
        _lib.make_problem(1, 1, 1, (4, 4, 4, 4), (2, 2, 2, 2), 0)


def test_status_translation(lib):
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.TNMF_EUNSUPPORTED)
    with pytest.raises(ValueError):
        _lib.check(_lib.TNMF_EINVAL)
    with pytest.raises(RuntimeError):
        _lib.check(_lib.TNMF_ECUDA + 2)
    _lib.check(0)


def test_tiled_path_selection(lib):
    f32_2d = _lib.make_problem(64, 3, 16, (256, 256), (11, 11), _lib.TNMF_F32)
    f64_2d = _lib.make_problem(64, 3, 16, (256, 256), (11, 11), _lib.TNMF_F64)
    f32_3d = _lib.make_problem(2, 1, 2, (8, 8, 8), (3, 3, 3), _lib.TNMF_F32)
    assert lib.tnmf_uses_tiled_path(ctypes.byref(f32_2d)) == 1
    assert lib.tnmf_uses_tiled_path(ctypes.byref(f64_2d)) == 0
    assert lib.tnmf_uses_tiled_path(ctypes.byref(f32_3d)) == 0
    # operation by operation: the tensor-core H update where most of every MMA is useful work, TMA where the strides
    # allow it (H rows of 266 floats are not 16-byte multiples)
    fam = lambda p, op: lib.tnmf_kernel_family(ctypes.byref(p), op)
    assert fam(f32_2d, _lib.OP_GRADIENT_H) == _lib.PATHS['tc']
    assert fam(f32_2d, _lib.OP_RECONSTRUCT) == _lib.PATHS['tc']
    # the switches of the family choice travel in tnmf_problem.flags (nothing is read from the environment)
    f32_2d.flags = _lib.FLAG_NO_TC_RECON
    assert fam(f32_2d, _lib.OP_RECONSTRUCT) == _lib.PATHS['tiled']
    f32_2d.flags = 0
    padded = _lib.make_problem(64, 3, 16, (256, 256), (11, 11), _lib.TNMF_F32, h_pitch=268)
    padded.flags = _lib.FLAG_NO_TC_RECON
    assert [fam(padded, op) for op in (0, 1, 2)] == [_lib.PATHS['tma'], _lib.PATHS['tc'], _lib.PATHS['tc']]
    padded.flags = 0
    assert [fam(padded, op) for op in (0, 1, 2)] == [_lib.PATHS['tc']] * 3
    padded.flags = _lib.FLAG_NO_TC_GRADW
    assert [fam(padded, op) for op in (0, 1, 2)] == [_lib.PATHS['tc'], _lib.PATHS['tc'], _lib.PATHS['tma']]
    padded.flags = _lib.FLAG_NO_TC
    assert [fam(padded, op) for op in (0, 1, 2)] == [_lib.PATHS['tma']] * 3
    padded.flags = _lib.FLAG_NO_TC | _lib.FLAG_NO_TMA
    assert [fam(padded, op) for op in (0, 1, 2)] == [_lib.PATHS['tiled']] * 3
    padded.flags = 0
    padded.reserved = 1
    assert fam(padded, 0) == -1
    padded.reserved = 0
    few_atoms = _lib.make_problem(8, 1, 3, (64, 64), (5, 5), _lib.TNMF_F32)       # 3 of 16 atoms, K 5 of 8: FP32 kernels
    assert fam(few_atoms, _lib.OP_GRADIENT_H) == _lib.PATHS['tma']
    tall_atom = _lib.make_problem(8, 2, 16, (64, 64), (16, 7), _lib.TNMF_F32)    # 16 atom rows do not fit the TMEM ring
    assert fam(tall_atom, _lib.OP_GRADIENT_H) == _lib.PATHS['tma']
    tall_atom.path = _lib.PATHS['tc']
    assert fam(tall_atom, _lib.OP_GRADIENT_H) == _lib.PATHS['tma']               # 'tc' = tensor cores where a kernel exists
    circ = _lib.make_problem(64, 3, 16, (256, 256), (11, 11), _lib.TNMF_F32, 'circular')
    assert [fam(circ, op) for op in (0, 1, 2)] == [_lib.PATHS['tiled']] * 3
    # single-channel 1-D batches run as one 2-D image of signal rows: TMA wherever the strides are 16-byte multiples
    # (V and R rows of 1000 floats are; H rows of 1049 floats are not until the backend pads them to 1052)
    one_d = _lib.make_problem(100, 1, 5, (1000,), (50,), _lib.TNMF_F32)
    assert [fam(one_d, op) for op in (0, 1, 2)] == [_lib.PATHS['tiled']] * 3          # below 2^20 elements: 1-D kernels
    big = _lib.make_problem(2048, 1, 64, (4096,), (128,), _lib.TNMF_F32, 'valid', 'auto', 64 * 4224, 4224)   # cfg4
    assert [fam(big, op) for op in (0, 1, 2)] == [_lib.PATHS['tma']] * 3
    one_d.flags = _lib.FLAG_ROWS_VIEW_ALWAYS
    assert [fam(one_d, op) for op in (0, 1, 2)] == [_lib.PATHS['tiled'], _lib.PATHS['tma'], _lib.PATHS['tiled']]
    one_d_padded = _lib.make_problem(100, 1, 5, (1000,), (50,), _lib.TNMF_F32, 'valid', 'auto', 5 * 1052, 1052,
                                     flags=_lib.FLAG_ROWS_VIEW_ALWAYS)
    assert [fam(one_d_padded, op) for op in (0, 1, 2)] == [_lib.PATHS['tma']] * 3
    one_d_padded.flags = _lib.FLAG_NO_ROWS_VIEW
    assert [fam(one_d_padded, op) for op in (0, 1, 2)] == [_lib.PATHS['tiled']] * 3
    one_d_2ch = _lib.make_problem(100, 2, 5, (1000,), (50,), _lib.TNMF_F32)
    assert [fam(one_d_2ch, op) for op in (0, 1, 2)] == [_lib.PATHS['tiled']] * 3
    one_d_circ = _lib.make_problem(100, 1, 5, (1000,), (50,), _lib.TNMF_F32, 'circular')
    assert [fam(one_d_circ, op) for op in (0, 1, 2)] == [_lib.PATHS['tiled']] * 3
    assert fam(f64_2d, _lib.OP_GRADIENT_W) == _lib.PATHS['generic']
    padded.path = _lib.PATHS['tma']
    assert fam(padded, 0) == _lib.PATHS['tma']
    one_d.path = _lib.PATHS['tma']
    assert fam(one_d, 0) == -1
    assert lib.tnmf_workspace_bytes(ctypes.byref(padded)) % 256 == 0
    bad_pitch = _lib.make_problem(2, 1, 2, (16, 16), (3, 3), _lib.TNMF_F32, h_pitch=10)    # narrower than a row
    assert fam(bad_pitch, 0) == -1
    f32_2d.path = _lib.PATHS['generic']
    assert lib.tnmf_uses_tiled_path(ctypes.byref(f32_2d)) == 0


def test_backend_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from tnmf_b200 import B200_Backend
    with pytest.raises(RuntimeError):
        B200_Backend()


def test_kernel_names_at_the_baseline_configurations(lib):
    """Which __global__ function serves each operation of the BASELINE geometries (tnmf_kernel_name; the planning code
    runs without a device): cfg2 and cfg3 on the tcgen05 kernels whose streamed operand lives in tensor memory, cfg4 (as
    one image of signal rows) and cfg5 on the FP32 kernels."""
    def names(p):
        return [lib.tnmf_kernel_name(ctypes.byref(p), op).decode() for op in (0, 1, 2)]
    cfg2 = _lib.make_problem(64, 3, 16, (256, 256), (11, 11), _lib.TNMF_F32, 'valid', 'auto', 16 * 266 * 268, 266 * 268, 268)
    assert names(cfg2) == ['recon_ts_kernel', 'hupd_ts_kernel', 'gradw_ts_kernel']
    cfg3 = _lib.make_problem(1024, 1, 32, (128, 128), (15, 15), _lib.TNMF_F32, 'valid', 'auto', 32 * 142 * 144, 142 * 144, 144)
    assert names(cfg3) == ['recon_os_kernel', 'hupd_ts_kernel', 'gradw_ns_kernel']
    cfg3.flags = _lib.FLAG_NO_TMEM_OPERAND                   # the round-1 forms (operands in shared memory)
    assert names(cfg3)[1:] == ['hupd_tc_kernel', 'gradw_tc_kernel']
    cfg3.flags = _lib.FLAG_NO_TC
    assert names(cfg3) == ['recon_tma_kernel', 'hupd_tma_kernel', 'gradw_tma_kernel']
    cfg4 = _lib.make_problem(2048, 1, 64, (4096,), (128,), _lib.TNMF_F32, 'valid', 'auto', 64 * 4224, 4224)
    assert names(cfg4) == ['recon_tma_kernel', 'hupd_tma_kernel', 'gradw_tma_kernel']
    cfg5 = _lib.make_problem(16, 1, 8, (512, 512), (64, 64), _lib.TNMF_F32, 'valid', 'auto', 8 * 575 * 576, 575 * 576, 576)
    assert names(cfg5) == ['recon_tma_kernel', 'tiled::hupd_kernel', 'gradw_tma_kernel']
    f64 = _lib.make_problem(4, 2, 3, (20, 17), (5, 3), _lib.TNMF_F64)
    assert names(f64) == ['generic_reconstruct_kernel', 'generic_gradient_h_kernel', 'generic_gradient_w_kernel']
    cfg2.reserved = 1
    assert names(cfg2) == ['none'] * 3


def test_peer_world_layout_and_buffer_size(lib):
    """struct tnmf_peer_world (include/tnmf_b200.h) and the size of the NVLink exchange buffer of
    tnmf_allreduce_update_w: data[2 parities][world][2 * count] + flags[world][atoms x channels]."""
    assert ctypes.sizeof(_lib.PeerWorld) == 8 + _lib.MAX_PEERS * ctypes.sizeof(ctypes.c_void_p)
    assert _lib.PeerWorld.buffers.offset == 8
    header = open(os.path.join(os.path.dirname(__file__), '..', 'include', 'tnmf_b200.h')).read()
    assert f'#define TNMF_MAX_PEERS {_lib.MAX_PEERS}' in header
    p = _lib.make_problem(0, 3, 16, (256, 256), (11, 11), _lib.TNMF_F32)
    count = 16 * 3 * 121
    for world in (1, 2, 8, 16):
        data = 2 * world * 2 * count * 4
        assert lib.tnmf_peer_buffer_bytes(ctypes.byref(p), world) == (data + 127) // 128 * 128 + world * 16 * 3 * 4
    assert lib.tnmf_peer_buffer_bytes(ctypes.byref(p), 17) == 0
    assert lib.tnmf_peer_buffer_bytes(ctypes.byref(p), 0) == 0
    # argument validation happens before anything is launched
    pw = _lib.PeerWorld()
    pw.world, pw.rank = 2, 0
    assert lib.tnmf_allreduce_update_w(ctypes.byref(p), 1, 1, ctypes.byref(pw), 1, 1e-9, None) == _lib.TNMF_EINVAL   # no buffers
    pw.world, pw.rank = 2, 2
    assert lib.tnmf_allreduce_update_w(ctypes.byref(p), 1, 1, ctypes.byref(pw), 1, 1e-9, None) == _lib.TNMF_EINVAL
