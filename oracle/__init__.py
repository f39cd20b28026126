"""CPU oracle for the xmhw threshold/detect hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker (or as the timed
CPU stand-in for the reference), never as a compute path of ``xmhw_b200``.

Contents
--------
xmhw_oracle.py   numpy restatement of the reference algorithm (float64), each
                 function citing the reference file:line it follows.
ref_harness.py   loads the UNMODIFIED reference modules from /root/reference
                 under stub ``xarray``/``dask`` modules (pins the oracle; only
                 usable where /root/reference exists, i.e. the build container).
nc_reader.py     stdlib HDF5 walker for the reference's NetCDF test data.
make_golden.py   writes tests/golden/*.npz from the reference's data + code.

Parity status: PINNED.  The restatement is checked against every golden vector
the reference's own tests hold for this path (tests/test_oracle_golden.py) and
fuzzed against the reference's pandas code run in the build container
(tests/test_oracle_vs_reference.py, skipped where /root/reference is absent).
"""
