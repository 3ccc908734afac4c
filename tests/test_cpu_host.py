"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/pybpl_b200.h declares,
host-side basis tables, the workload generator, and the multi-rank combine logic under gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as ge
    ge.build()
    from pybpl_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "pybpl_b200.h")).read()
    declared = set(re.findall(r"\b(bpl_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTS) <= declared
    assert L.bpl_version() == 1
    ps = _lib.default_params()
    assert (ps.imsize, ps.fsize, ps.ink_ncon) == (105, 11, 2) and ps.ink_b == 6.0


def test_sass_is_sm100a():
    from pybpl_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out


def test_host_basis_tables_close_to_reference():
    from pybpl_b200 import ops
    from tests.util_golden import load
    d = load("splines")
    assert np.array_equal(ops._basis_host(5, 200), d["table_5_200"])       # shipped reference table
    for nland, neval in d["table_pairs"]:
        C = ops._basis_host(int(nland), int(neval))
        assert np.abs(C - d[f"table_{nland}_{neval}"]).max() < 1e-6
    import torch
    C = ops.basis_rows(5, torch.from_numpy(d["s_explicit"])).numpy()
    assert np.abs(C - d["table_s_explicit"]).max() < 1e-6


def test_no_cpu_fallback():
    import torch
    from pybpl_b200 import _lib, ops
    with pytest.raises(_lib.BplError):
        ops.pack_images(torch.zeros(105, 105))
    with pytest.raises(_lib.BplError):
        ops.spline_eval(torch.zeros(200, 5), torch.zeros(1, 5, 2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pybpl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "bpl_oracle" not in src, f


def test_workload_generator_is_seeded_and_shardable():
    from pybpl_b200 import workload
    a = workload.synth_tokens(512, seed=3)
    b = workload.synth_tokens(512, seed=3)
    assert all(np.array_equal(a[k], b[k]) for k in a if isinstance(a[k], np.ndarray))
    tso = a["tok_stroke_off"]; sso = a["stroke_sub_off"]
    K = np.diff(tso); S = np.diff(sso[tso])
    assert K.min() >= 1 and K.max() <= 5 and S.max() <= 50 and (a["invscale"] > 0).all()
    parts = [workload.shard(a, r, 3) for r in range(3)]
    assert sum(p["P"] for p in parts) == 512
    assert np.array_equal(np.concatenate([p["ctrl"] for p in parts]), a["ctrl"])
    assert np.array_equal(np.concatenate([p["sigma"] for p in parts]), a["sigma"])


def test_oracle_runs_workload():
    from oracle import Oracle
    from pybpl_b200 import workload
    w = workload.synth_tokens(16, seed=11)
    o = Oracle(np.float32)
    C = np.load(os.path.join(ROOT, "pybpl_b200", "data", "basis_5_200.npy"))
    import oracle as _oracle
    img = _oracle.target_image_for(w, o, C)
    r = o.token_batch(C, w["tok_stroke_off"], w["stroke_sub_off"], w["ctrl"], w["invscale"], w["position"], w["eps"],
                      w["sigma"], img[None])
    assert np.isfinite(r["ll"]).all() and (r["ll"] < 0).all()


_GLOO = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from pybpl_b200 import sweep
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = torch.Generator().manual_seed(0)
ll = -50 - 2000 * torch.rand(1001, generator=g)
per = (1001 + world - 1) // world
mine = ll[rank * per:(rank + 1) * per]
part = sweep.partial_from_scores(mine)
tot = sweep.combine_partials(part, group=None)
want = torch.logsumexp(ll.double(), 0).item()
assert abs(tot - want) < 1e-6 * abs(want), (tot, want)
# empty shard
tot2 = sweep.combine_partials(sweep.partial_from_scores(mine if rank == 0 else mine[:0]), group=None)
want2 = torch.logsumexp(ll[:per].double(), 0).item()
assert abs(tot2 - want2) < 1e-6 * abs(want2)
# cross-scoring (configs[4]): contiguous row shards of a [tokens, images] matrix, uneven (1001 = 501 + 500), one gather
full = torch.arange(1001 * 7, dtype=torch.float32).view(1001, 7)
lo, hi = sweep.shard_bounds(1001, rank, world)
assert (lo, hi) == ((0, 501) if rank == 0 else (501, 1001))
got = sweep.gather_rows(full[lo:hi].clone(), 1001)
assert got.shape == full.shape and torch.equal(got, full)
assert sum(sweep.shard_bounds(10, r, 4)[1] - sweep.shard_bounds(10, r, 4)[0] for r in range(4)) == 10
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sweep_combine_world2_gloo(tmp_path):
    script = tmp_path / "gloo_sweep.py"
    script.write_text(_GLOO.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.count("ok") >= 2


def test_bench_reference_arm_runs():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0
    # the reference's own torch code when baseline/_ref is installed (kind "reference"), else the C oracle port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline_port"]["kind"] == "port"
    assert line["value"] == line["cpu_baseline"]["value"] == line["e2e"]["value"]


def test_dropin_rebinds_reference_call_sites():
    """dropin.install() must rebind every hot-path name of the reference and uninstall() restore it.
    Needs the reference checkout (build container only); no compute is run."""
    ref = os.environ.get("PYBPL_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "pybpl")):
        pytest.skip("reference checkout not present")
    code = r'''
import sys, types
sys.dont_write_bytecode = True
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, %r); sys.path.insert(0, %r)
import pybpl.rendering as R, pybpl.splines as S, pybpl.objects.part as P, pybpl.objects.relation as REL
import pybpl.model.image_dist as ID
from pybpl_b200 import dropin, rendering, splines, objects, model
orig = (R.render_image, ID.render_image, S.get_stk_from_bspline, REL.get_stk_from_bspline, P.vanilla_to_motor,
        ID.CharacterImageDist.get_pimg, ID.CharacterImageDist.score_image, ID.CharacterImageDist.sample_image)
dropin.install()
assert R.render_image is rendering.render_image and ID.render_image is rendering.render_image
assert S.get_stk_from_bspline is splines.get_stk_from_bspline and REL.get_stk_from_bspline is splines.get_stk_from_bspline
assert S.bspline_eval is splines.bspline_eval
assert P.vanilla_to_motor is objects.vanilla_to_motor
assert ID.CharacterImageDist.get_pimg is model.CharacterImageDist.get_pimg
assert ID.CharacterImageDist.score_image is model.CharacterImageDist.score_image
assert ID.CharacterImageDist.sample_image is model.CharacterImageDist.sample_image
dropin.uninstall()
now = (R.render_image, ID.render_image, S.get_stk_from_bspline, REL.get_stk_from_bspline, P.vanilla_to_motor,
       ID.CharacterImageDist.get_pimg, ID.CharacterImageDist.score_image, ID.CharacterImageDist.sample_image)
assert all(a is b for a, b in zip(orig, now)) and not hasattr(S, "bspline_eval")
print("dropin ok")
''' % (ref, ROOT)
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=240)
    assert r.returncode == 0 and "dropin ok" in r.stdout, r.stderr[-2000:]


def test_parse_file_roundtrip_and_layout(tmp_path):
    """pybpl_b200/packed.py: byte-exact round trip, the header layout include/pybpl_parsefile.h states,
    bit packing of images, rejection of corrupt files."""
    import struct
    import numpy as np
    from pybpl_b200 import packed, workload
    w = workload.synth_tokens(300, seed=3, sigma_mode="uniform")
    ty = workload.synth_types(w, 4)
    rng = np.random.default_rng(0)
    imgs = (rng.random((7, 105, 105)) < 0.1).astype(np.uint8)
    path = str(tmp_path / "parses.bpl")
    n = packed.save(path, w, ty, imgs)
    assert n == os.path.getsize(path) and n % 16 == 0
    raw = open(path, "rb").read()
    magic, hb, nsec, P, K, S, T, ncpt, neval = struct.unpack_from("<8sII6i", raw, 0)
    assert magic == b"BPLPARS1" and hb % 16 == 0 and (P, T, ncpt, neval) == (300, 7, 5, 200)
    assert nsec == 7 + 8 + 1 and K == len(w["stroke_sub_off"]) - 1 and S == w["ctrl"].shape[0]
    pf = packed.ParseFile.load(path, pin=False)
    a = pf.arrays()
    for k in packed.TOKEN_KEYS:
        assert np.array_equal(a[k], w[k]) and a[k].dtype == w[k].dtype
    for k in packed.TYPE_KEYS:
        assert np.array_equal(a[k], ty[k])
    assert np.array_equal(packed.unpack_bits(a["image_bits"]), imgs)
    i = 105 * 17 + 33                                    # pixel (17, 33) of image 2 = bit i&31 of word i>>5
    assert ((int(a["image_bits"][2, i >> 5]) >> (i & 31)) & 1) == imgs[2, 17, 33]
    assert packed.to_bytes(a | {"neval": 200}, {k: a[k] for k in packed.TYPE_KEYS}, image_bits=a["image_bits"]) == raw
    with pytest.raises(ValueError):
        packed.ParseFile(b"NOTAFILE" + raw[8:], pin=False)
    bad = bytearray(raw); struct.pack_into("<i", bad, hb + 4, 10 ** 6)       # tok_stroke_off[1] out of range
    with pytest.raises(ValueError):
        packed.ParseFile(bytes(bad), pin=False)
    with pytest.raises(ValueError):
        packed.pack_bits(np.full((1, 105, 105), 0.5))


def test_parse_file_rejects_malformed_sections(tmp_path):
    """ParseFile checks every section against the header's counts, the attachment indices against the token structure and
    the images' padding bits: the kernels index these arrays without bounds checks."""
    import numpy as np
    from pybpl_b200 import packed, workload
    w = workload.synth_tokens(120, seed=4, sigma_mode="uniform")
    ty = workload.synth_types(w, 5)
    imgs = (np.random.default_rng(1).random((3, 105, 105)) < 0.1).astype(np.uint8)
    bits = packed.pack_bits(imgs)
    packed.ParseFile(packed.to_bytes(w, ty, imgs), pin=False)                        # the well-formed file loads

    def rejected(w2=w, ty2=ty, bits2=bits):
        with pytest.raises(ValueError):
            packed.ParseFile(packed.to_bytes(w2, ty2, image_bits=bits2), pin=False)
    rejected(w2=dict(w, eps=w["eps"][:-1]))                                          # eps shorter than P
    rejected(w2=dict(w, position=np.concatenate([w["position"], w["position"][:1]])))  # position longer than K
    later = np.flatnonzero(ty["rel_cat"] != 0)
    assert later.size
    bad = dict(ty, attach_ix=ty["attach_ix"].copy()); bad["attach_ix"][later[0]] = 31
    rejected(ty2=bad)                                                                # names a stroke that does not exist
    bad = dict(ty, rel_cat=ty["rel_cat"].copy()); bad["rel_cat"][0] = 7
    rejected(ty2=bad)
    mids = np.flatnonzero(ty["rel_cat"] == 3)
    if mids.size:
        bad = dict(ty, attach_subix=ty["attach_subix"].copy()); bad["attach_subix"][mids[0]] = 99
        rejected(ty2=bad)
    b2 = bits.copy(); b2[1, -1] |= np.uint32(1 << 20)                                # a padding bit
    rejected(bits2=b2)


def test_host_stager_packs_the_same_bytes_with_any_number_of_threads():
    """ops.HostSweepStager: the ragged packing of a chunk (offsets rebased, the chunk's slices copied into one staging
    buffer) gives the same sections whether the control-point copy runs on one thread or is split over a pool."""
    import numpy as np
    from pybpl_b200 import ops, workload
    w = workload.synth_tokens(20000, seed=3, sigma_mode="uniform")
    a = ops.HostSweepStager(w, "cpu", chunk=8192, pin=False, threads=1)
    b = ops.HostSweepStager(w, "cpu", chunk=8192, pin=False, threads=4)
    assert len(a.chunks) == 3 and b.threads == 4
    for c in range(len(a.chunks)):
        a.pack(c, 0); b.pack(c, 1)
        c0, c1, k0, k1, s0, s1 = a.chunks[c]
        rows = dict(tok_stroke_off=c1 - c0 + 1, stroke_sub_off=k1 - k0 + 1, ctrl=s1 - s0, invscale=s1 - s0, position=k1 - k0,
                    eps=c1 - c0, sigma=c1 - c0)
        for name, n in rows.items():
            assert np.array_equal(a._host_view(0, name, n), b._host_view(1, name, n)), (c, name)
        assert np.array_equal(b._host_view(1, "ctrl", s1 - s0), w["ctrl"][s0:s1].reshape(-1, 10))
        assert b._host_view(1, "tok_stroke_off", c1 - c0 + 1)[0] == 0 and b._host_view(1, "stroke_sub_off", k1 - k0 + 1)[0] == 0
