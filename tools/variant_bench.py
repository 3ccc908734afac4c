#!/usr/bin/env python
"""Build the library with different -D flags on the GPU box and time each build with tools/ab_forward.py:
    python tools/variant_bench.py "name:ENV=1,ENV2=x:-DFOO=1 -DBAR=2" ...
(name, environment assignments and nvcc flags separated by ':'; empty parts allowed)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = [os.path.join(ROOT, "pybpl_b200", "csrc", f) for f in ("bpl_render.cu", "bpl_splines.cu", "bpl_token_prior.cu", "bpl_token_sample.cu", "bpl_type_sample.cu", "bpl_type_prior.cu")]
procs = []
for i, spec in enumerate(sys.argv[1:]):
    name, envs, flags = (spec.split(":") + ["", ""])[:3]
    so = f"/tmp/libbpl_var_{i}.so"
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-w",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"] + flags.split() + src + ["-o", so]
    procs.append((name, envs, so, subprocess.Popen(cmd)))
for name, envs, so, pr in procs:
    if pr.wait() != 0:
        print(f"{name}: BUILD FAILED"); continue
    env = dict(os.environ, BPL_LIB=so)
    for kv in filter(None, envs.split(",")):
        k, v = kv.split("=", 1); env[k] = v
    sys.stdout.flush()
    subprocess.call([sys.executable, os.path.join(ROOT, "tools", os.environ.get("VARIANT_SCRIPT", "ab_forward.py")), name] + os.environ.get("VARIANT_ARGS", "").split(), env=env)
