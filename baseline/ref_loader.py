"""Loads the UNMODIFIED reference (rfeinman/pyBPL) installed under baseline/_ref by baseline/install_ref.sh.

Used only by the parity tests (the drop-in is exercised over the real reference objects) and by bench.py's
`cpu_baseline_torch` / `--impl reference` legs (the reference's own torch code timed on the host cores).  The product
(pybpl_b200/) never imports this.

Two accommodations, both outside the reference's files:
  * `matplotlib` is not installed in this image and pybpl/library/spatial_OLD imports it at module level -> an empty
    stub module is put into sys.modules first;
  * `grad_patch=True` re-executes pybpl/rendering.py in memory with the in-place `dist.clamp_` of line 97 made
    out-of-place (forward bit-identical; torch >= 2 refuses to differentiate through the original).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "pybpl", "__init__.py")) and os.path.isdir(os.path.join(REF, "lib_data"))


def load(grad_patch=False):
    """Returns the imported `pybpl` package (raises ImportError when baseline/_ref is absent)."""
    if not available():
        raise ImportError("baseline/_ref is missing: run baseline/install_ref.sh in the build container")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import pybpl
    import pybpl.library  # noqa: F401
    import pybpl.model  # noqa: F401
    if grad_patch and not getattr(pybpl, "_b200_grad_patch", False):
        import pybpl.rendering as R
        import pybpl.model.image_dist as ID
        src = open(R.__file__).read()
        bad = "dist.clamp_(None, ps.ink_max_dist)"
        assert bad in src
        exec(compile(src.replace(bad, "dist = dist.clamp(None, ps.ink_max_dist)"), R.__file__, "exec"), R.__dict__)
        ID.render_image = R.render_image
        pybpl._b200_grad_patch = True
    return pybpl
