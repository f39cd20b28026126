"""Recipe that makes the reference's OWN detect code available where /root/reference is not
(the GPU box): copies the three pure-pandas / numpy modules of coecms/xmhw v0.9.3

    xmhw/exception.py   xmhw/features.py   xmhw/identify.py

unmodified from the read-only checkout into oracle/_ref/xmhw/.  oracle/_ref/ is git-ignored (no
reference source enters the history) but not gpurun-ignored, so it travels with the snapshot like
a built .so.  TEST / BASELINE INFRASTRUCTURE ONLY: oracle/ref_harness.py loads the copies under
stub `xarray` / `dask` modules for tests/test_oracle_vs_reference.py and for the CPU arm of
bench.py (`--impl reference`, kind "reference-pandas+glue-port"); the product never imports them.

    python -m oracle.build_ref        (run by __graft_entry__.build() when /root/reference exists)
"""
import filecmp
import os
import shutil

SRC = os.environ.get("XMHW_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("exception.py", "features.py", "identify.py")


def build():
    """Returns True when oracle/_ref holds the three files (copied now or earlier)."""
    have_src = all(os.path.isfile(os.path.join(SRC, "xmhw", f)) for f in FILES)
    if have_src:
        os.makedirs(os.path.join(DST, "xmhw"), exist_ok=True)
        for f in FILES:
            s, d = os.path.join(SRC, "xmhw", f), os.path.join(DST, "xmhw", f)
            if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
    return all(os.path.isfile(os.path.join(DST, "xmhw", f)) for f in FILES)


if __name__ == "__main__":
    print("oracle/_ref ready:", build())
