"""CPU oracle for the multi-scale view-synthesis loss.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline.  The product (the package next to this
directory) never imports ``oracle`` and fails loudly when ``libtdl.so`` is
missing.

Contents
--------
``restatement``  a plain-PyTorch (CPU, fp32 or fp64) restatement of the
                 reference's loss path, same operation order, each function
                 citing the reference file:line it follows.
``ref_loader``   imports the *real* reference classes from ``/root/reference``
                 (only exists in the build container) to pin the restatement
                 and to generate ``tests/golden/*.pt``.

Parity pin: the reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the pin is "the reference's own Python executed by
this container's torch 2.11 in fp32 on seeded synthetic inputs"; those outputs
are committed under ``tests/golden/`` together with ``make_golden.py``.
"""
