"""CPU-only tests of the host side: the C ABI library loads and exports every symbol the header declares (no
compute without a GPU), the product binding refuses to run without CUDA, label tables of the nnUNet path, and the
pair-parallel sharding over a world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    from convexadam_b200 import _build
    return _build.build()


def test_header_symbols_exported(so_path):
    hdr = open(os.path.join(ROOT, "include", "convexadam_b200.h")).read()
    declared = set(re.findall(r"^CA_API [\w \*]+?\b(ca_\w+)\(", hdr, flags=re.M))
    assert len(declared) >= 19
    dll = ctypes.CDLL(so_path)
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in the header but not exported"
    from convexadam_b200 import _native
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    dll.ca_abi_version.restype = ctypes.c_int
    assert dll.ca_abi_version() == _native.ABI_VERSION
    assert dll.ca_is_cuda_build() == 1


def test_library_is_sm100a_only(so_path):
    out = subprocess.run(["cuobjdump", "-lelf", so_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_cuda():
    from convexadam_b200._native import NativeError, get_lib
    from convexadam_b200.convex_adam_MIND import convex_adam_pt
    with pytest.raises(NativeError):
        get_lib()
    with pytest.raises(RuntimeError):
        convex_adam_pt(np.zeros((8, 8, 8)), np.zeros((8, 8, 8)))


def test_product_never_imports_oracle_or_emulation():
    pkg = os.path.join(ROOT, "convexadam_b200")
    bad = re.compile(r"^\s*(from|import)\s+(oracle|tests)\b|ca_oracle|libca_oracle|libca_emu", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                assert not bad.search(open(os.path.join(dp, f)).read()), f"{f} reaches into the oracle / test build"


def test_validate_image():
    from convexadam_b200.convex_adam_utils import validate_image
    t = validate_image(np.ones((2, 3, 4), np.int16))
    assert isinstance(t, torch.Tensor) and t.dtype == torch.float64
    x = torch.zeros(2, 2, 2)
    assert validate_image(x) is x
    with pytest.raises(ValueError):
        validate_image([1, 2, 3])


def test_nnunet_label_tables_match_reference_formula():
    from convexadam_b200.convex_adam_nnUNet import _label_tables
    g = torch.Generator().manual_seed(0)
    a = torch.randint(0, 7, (6, 6, 6), generator=g).float()
    b = torch.randint(0, 7, (6, 6, 6), generator=g).float()
    a[a == 4] = 2
    b[b == 4] = 2
    pos, weight, lut = _label_tables(a, b)
    assert pos.tolist() == [0, 1, 2, 3, 5, 6] and lut.tolist() == [0, 1, 2, 3, -1, 4, 5]
    cnt = torch.bincount(a.long().reshape(-1), minlength=7) + torch.bincount(b.long().reshape(-1), minlength=7)
    w = 1 / (cnt[pos] + 1e-32).float().pow(0.3)
    assert torch.equal(weight, w / w.mean())


def test_shard_indices():
    from convexadam_b200.batch import shard_indices
    assert shard_indices(32, 3, 8) == [3, 11, 19, 27]
    assert sorted(sum((shard_indices(10, r, 4) for r in range(4)), [])) == list(range(10))
    assert shard_indices(2, 5, 8) == []


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from convexadam_b200.batch import register_pairs
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
calls = []
def fake(fixed, moving, **kw):
    calls.append(int(fixed[0, 0, 0]))
    return np.full((2, 2, 2, 3), float(fixed[0, 0, 0]) + kw["bias"])
pairs = [(torch.full((2, 2, 2), float(i)), torch.zeros(2, 2, 2)) for i in range(5)]
res = register_pairs(pairs, gather=True, register_fn=fake, bias=0.5)
rank = dist.get_rank()
assert calls == list(range(rank, 5, 2)), calls
if rank == 0:
    assert all(r is not None and float(r[0, 0, 0, 0]) == i + 0.5 for i, r in enumerate(res)), res
else:
    assert [r is not None for r in res] == [False, True, False, True, False]
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_pair_parallel_gloo_world2(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) on a tiny shape: one JSON line with the
    contract's keys, produced without touching CUDA."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--shape", "24", "20", "32", "--niter", "3"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0
    from oracle import ref_shim
    if ref_shim.available():      # the reference's own code (here: /root/reference; on the GPU box: the staged oracle/_ref)
        assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["port"]["value"] > 0
    else:
        assert line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_translation_helpers_match_reference_formula():
    """convex_adam_translation.py:12-29 and :99-104 restated: rounding to whole voxels of the moving spacing, (x,y,z) order."""
    import numpy as np
    from convexadam_b200.convex_adam_translation import field_to_translation, index_translation_to_world_translation
    from oracle import ca_oracle as orc
    rs = np.random.RandomState(0)
    for _ in range(100):
        t, sp = rs.randn(3) * 7, rs.rand(3) * 3 + 0.2
        field = np.broadcast_to(t, (2, 3, 2, 3)).copy()
        mean, want = orc.translation_from_field(field, sp)
        assert np.allclose(mean, t) and field_to_translation(t, sp) == want
    assert field_to_translation((2.5, -0.5, 1.49), (1.0, 1.0, 1.0)) == (1.0, -0.0, 2.0)        # half to even, reversed
    got = index_translation_to_world_translation((1, 2, 3), (0, 1, 0, -1, 0, 0, 0, 0, 1))
    assert np.allclose(got, [2.0, -1.0, 3.0])


def test_overlap_setting_parses(monkeypatch):
    """Params.overlap / CONVEXADAM_B200_OVERLAP: 0 | 1 | auto (default), anything else is an error; the explicit field wins."""
    from convexadam_b200.pipeline import Params
    monkeypatch.delenv("CONVEXADAM_B200_OVERLAP", raising=False)
    assert Params().resolved_overlap() is None
    assert Params(overlap="1").resolved_overlap() is True and Params(overlap="0").resolved_overlap() is False
    monkeypatch.setenv("CONVEXADAM_B200_OVERLAP", "0")
    assert Params().resolved_overlap() is False and Params(overlap="1").resolved_overlap() is True
    monkeypatch.setenv("CONVEXADAM_B200_OVERLAP", "auto")
    assert Params().resolved_overlap() is None
    monkeypatch.setenv("CONVEXADAM_B200_OVERLAP", "sometimes")
    with pytest.raises(ValueError):
        Params().resolved_overlap()


def test_kovesi_widths_match_reference():
    """search.kovesi_widths against the pool widths the reference's own kovesi_spline(sigma, 4) builds (values produced in this
    container by executing that function from /root/reference/self_configuring/convexAdam_hyper_util.py:475-488), for every sigma
    of the search's `avgs` list (adam_run_withconfig_shiftSpline.py:140-141) and two more."""
    from convexadam_b200.search import kovesi_widths
    table = {0.7: [3], 1.0: [3, 3], 1.3: [3, 3, 3], 1.6: [3, 3, 3, 3], 1.9: [3, 3, 3, 5], 2.2: [3, 3, 5, 5], 2.5: [3, 5, 5, 5],
             2.8: [5, 5, 5, 5], 3.5: [5, 5, 7, 7]}
    for sigma, want in table.items():
        assert kovesi_widths(sigma, 4) == want, (sigma, kovesi_widths(sigma, 4))


def test_gaussian_weights_match_reference():
    """search.gaussian_weights against the weights the reference's own GaussianSmoothing(sigma) holds (bit patterns produced in this
    container by executing that class from /root/reference/self_configuring/convexAdam_hyper_util.py:454-464)."""
    from convexadam_b200.search import gaussian_weights
    table = {0.7: ['0x1.3b3ae40000000p-7', '0x1.a4b52e0000000p-3', '0x1.23cb920000000p-1', '0x1.a4b52e0000000p-3', '0x1.3b3ae40000000p-7'],
             1.0: ['0x1.be5f100000000p-5', '0x1.f41fd80000000p-3', '0x1.9c48680000000p-2', '0x1.f41fd80000000p-3', '0x1.be5f100000000p-5'],
             2.0: ['0x1.1f5f620000000p-4', '0x1.0c70fc0000000p-3', '0x1.8694700000000p-3', '0x1.ba95be0000000p-3', '0x1.8694700000000p-3',
                   '0x1.0c70fc0000000p-3', '0x1.1f5f620000000p-4']}
    for sigma, want in table.items():
        assert [float.hex(x) for x in gaussian_weights(sigma)] == want, sigma
