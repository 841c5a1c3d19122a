"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference (r9y9/gantts @ fb1e75f) GAN-step hot path, used as the
*checker* for the CUDA product path in ``gantts_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may
import anything from here.  Nothing under ``gantts_b200/`` imports it, and the product path
fails loudly when ``libgantts_b200.so`` is missing -- there is no CPU fallback.

Parity pinning status (see DESIGN.md "Oracle"):
  * ``gantts_port``      -- PINNED against the reference's own Python (imported read-only from
                            /root/reference in the build container by tests/golden/make_golden.py;
                            the outputs are committed as tests/golden/*.npz).
  * ``nnmnkwii_port``    -- the MLPG arithmetic lives in the un-vendored third-party package
                            nnmnkwii (>= 0.0.14, reference setup.py:63,66) which is absent from
                            /root/reference and from this image: **parity unpinned** against
                            nnmnkwii itself; restated from its published definition
                            R = (W^T W)^-1 W^T and anchored on the reference's call sites and
                            on the reference's five unit tests, which all pass on top of it.
  * SRU (``sru_port``)   -- third-party github.com/taolei87/sru, not vendored, no reference test
                            touches it: **parity unpinned**.
"""
