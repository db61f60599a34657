"""CPU-side checks of the boundary: the library loads without a GPU and exports every declared symbol."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from physdock_b200 import build, _lib
    build.build_library()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "physdock_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(lib):
    from physdock_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 29
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/physdock_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"


def test_host_only_entry_points(lib):
    from physdock_b200 import _lib
    assert lib.pdk_abi_version() == _lib.ABI_VERSION == 2
    assert lib.pdk_pad_len(2048) == 2048 and lib.pdk_pad_len(1325) == 1408 and lib.pdk_pad_len(1) == 128


def test_handle_lifecycle_and_sizes_without_gpu(lib):
    import ctypes as C
    from physdock_b200 import _lib
    dims = _lib.DitDims(c_a=128, c_ap=16, c_s=512, c_z=128, n_atom_blocks=3, n_token_blocks=12, hidden_a=384,
                        hidden_s=1408, n_mod=41472, sigma_data=16.0, eps=1e-8, inf=1e9)
    h = C.c_void_p()
    assert lib.pdk_dit_create(C.byref(dims), C.byref(h)) == 0
    ab, tb, ws = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert lib.pdk_dit_bias_bytes(h, 2048, 256, C.byref(ab), C.byref(tb)) == 0
    assert ab.value == 6 * 4 * 2048 * 2048 * 4 and tb.value == 12 * 16 * 256 * 256 * 4     # 403 MB + 50 MB
    assert lib.pdk_dit_workspace_bytes(h, 16, 2048, 256, C.byref(ws)) == 0
    assert 100e6 < ws.value < 2e9
    assert lib.pdk_dit_launches_per_denoise(h) == 2 + 5 * 6 + 7 * 12 + 6     # atom transition fused; precond / gather-add ride in an AdaLN
    assert lib.pdk_dit_launches_per_denoise_cond(h) == 5 * 6 + 7 * 12 + 6     # conditioning hoisted out of the step
    assert lib.pdk_dit_cond_width(h) == 41472 + 8
    cb = C.c_size_t()
    assert lib.pdk_dit_conditioning_workspace_bytes(h, 40, C.byref(cb)) == 0 and cb.value == 2 * 128 * 256 * 2
    assert lib.pdk_dit_denoise_cond(h, None, None, 0, 1, None, 0, None, None, None) != 0
    assert b"no prepared complex" in lib.pdk_last_error()
    # errors are reported, not swallowed
    assert lib.pdk_dit_denoise(h, None, None, 1, None, 0, None, None) != 0
    assert b"no prepared complex" in lib.pdk_last_error()
    bad = _lib.DitDims(c_a=64, c_ap=16, c_s=512, c_z=128, n_atom_blocks=3, n_token_blocks=12, hidden_a=384,
                       hidden_s=1408, n_mod=41472, sigma_data=16.0, eps=1e-8, inf=1e9)
    h2 = C.c_void_p()
    assert lib.pdk_dit_create(C.byref(bad), C.byref(h2)) != 0
    assert lib.pdk_dit_destroy(h) == 0


def test_missing_library_fails_loudly(tmp_path):
    from physdock_b200 import _lib
    with pytest.raises(_lib.PdkError):
        _lib.load(str(tmp_path / "nope.so"))


def test_module_mirrors_reference_state_dict():
    from physdock_b200.dit import B200DiT
    from physdock_b200.synthetic import DiTDims, dit_param_shapes
    for name in ("toy", "medium"):
        d = DiTDims.named(name)
        m = B200DiT(no_blocks_atom=d.no_blocks_atom, no_blocks_dit=d.no_blocks_dit)
        sd = m.state_dict()
        assert list(sd.keys()) == list(dit_param_shapes(d).keys())
        assert all(tuple(sd[k].shape) == s for k, s in dit_param_shapes(d).items())
    assert sum(p.numel() for p in B200DiT().parameters()) == 50772160      # 50.77 M


@pytest.mark.parametrize("B,H,QT", [(16, 4, 16), (16, 16, 2), (1, 4, 1), (5, 4, 2), (7, 16, 1), (40, 4, 3), (48, 16, 3), (3, 4, 24)])
def test_attention_work_list_covers_every_sample_once(lib, B, H, QT):
    """Host logic of attention_umma.cu: balanced sample groups, largest first, every (head, q tile, sample) exactly once."""
    import ctypes as C
    buf = (C.c_uint32 * 800)()
    n = lib.pdk_attention_work_list(B, H, QT, 148, buf, 800)
    assert 0 < n <= 800
    seen = set()
    sizes = []
    for i in range(n):
        e = buf[i]
        h, qt, b0, ng = e & 0xff, (e >> 8) & 0xff, (e >> 16) & 0xff, e >> 24
        assert 1 <= ng <= 4 and h < H and qt < QT and b0 + ng <= B
        sizes.append(ng)
        for b in range(b0, b0 + ng):
            assert (h, qt, b) not in seen
            seen.add((h, qt, b))
    assert len(seen) == B * H * QT
    assert sizes == sorted(sizes, reverse=True)
    if (B, H, QT) == (16, 4, 16):       # the benchmark shape: one 4-sample and one 3-sample CTA per SM
        assert n == 292 and sizes.count(4) == 148 and sizes.count(3) == 144


def test_attention_work_list_limits(lib):
    import ctypes as C
    buf = (C.c_uint32 * 800)()
    assert lib.pdk_attention_work_list(64, 4, 24, 148, buf, 800) == -1      # 96 pairs x 17 groups: the library splits the samples
    assert lib.pdk_attention_work_list(16, 4, 16, 148, buf, 10) == -2
