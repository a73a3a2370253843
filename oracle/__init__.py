"""CPU oracle for the FEM-FCT-PDECO hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the reference's algorithm
(KarolinaBenkova/FEM-FCT-PDECO: helpers.py, old_helpers.py and the dolfin P1
assembly they call).  It exists to check the CUDA path; it is never the thing
shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package (``fem-fct-pdeco_b200/``) must not import anything from
here and fails loudly when its CUDA library is missing.

Parity pinning (see DESIGN.md "Oracle"): the restatement is checked against
 * the reference's shipped data files (chemotaxis 10-step trajectory,
   solid-body t=0.25/0.5/1 fields) -> tests/golden/*.npy, tests/test_oracle_golden.py
 * the reference's own FCT_alg_ref / ChebSI / artificial_diffusion_mat /
   L2_norm_sq_Q / cost_functional, imported unmodified from /root/reference
   with dolfin/matplotlib stubbed (oracle/ref_loader.py; only available in the
   build container) -> fixtures made by tests/golden/make_golden.py.
"""
