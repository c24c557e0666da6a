"""Import the UNMODIFIED reference (`/root/reference/kbbq`) in this container.

Test infrastructure only (golden-vector generation, see tests/golden/make_golden.py).  The
reference needs pysam/matplotlib/seaborn/khmer (absent here) and numpy aliases removed in
NumPy >= 1.24; this module supplies stand-ins for the former and re-adds the latter, then
imports the reference from where it lies.  Nothing under /root/reference is copied or modified,
and nothing on the GPU box may import this module (the reference does not travel).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("KBBQ_REFERENCE_ROOT", "/root/reference")


def _pandas_aliases(pd):
    """pandas API the reference's report code (kbbq/gatk/bqsr.py:227-366, kbbq/recaltable.py) relies on
    and pandas >= 2 removed: DataFrame.append, object-dtype strings, fillna(downcast='infer')."""
    import numpy as np
    if not hasattr(pd.DataFrame, "append"):
        pd.DataFrame.append = lambda self, other: pd.concat([self, other])
    try:
        pd.set_option("future.infer_string", False)
    except Exception:
        pass
    if not getattr(pd.Series.fillna, "_kbbq_shim", False):
        orig = pd.Series.fillna

        def fillna(self, value=None, downcast=None, **kw):
            out = orig(self, value, **kw)
            if downcast == "infer" and out.dtype.kind == "f" and len(out) and \
                    np.all(np.isfinite(out.values)) and np.all(out.values == np.floor(out.values)):
                out = out.astype(np.int64)
            return out
        fillna._kbbq_shim = True
        pd.Series.fillna = fillna


def load():
    """Return (recalibrate, compare_reads, applybqsr) modules of the reference."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "kbbq")):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    # these must be imported BEFORE the aliases are patched
    import scipy.stats  # noqa: F401
    import sklearn.linear_model  # noqa: F401
    import sklearn.isotonic  # noqa: F401
    import pandas  # noqa: F401
    for alias, target in (("int", int), ("float", float), ("unicode", np.str_),
                          ("NINF", -np.inf), ("object", object)):
        if not hasattr(np, alias):
            setattr(np, alias, target)
    _pandas_aliases(pandas)
    here = os.path.dirname(os.path.abspath(__file__))
    # a product package named `kbbq` may already be imported; the reference must win here
    for name in [m for m in sys.modules if m == "kbbq" or m.startswith("kbbq.")]:
        del sys.modules[name]
    sys.path.insert(0, os.path.join(here, "stubs"))
    sys.path.insert(0, REFERENCE_ROOT)
    import kbbq.recalibrate as recalibrate
    import kbbq.compare_reads as compare_reads
    from kbbq.gatk import applybqsr
    assert recalibrate.__file__.startswith(REFERENCE_ROOT)
    return recalibrate, compare_reads, applybqsr
