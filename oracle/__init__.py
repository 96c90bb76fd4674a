"""CPU oracle for the two-for-one-diffusion hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, in plain CPU PyTorch, the arithmetic the reference performs on the
hot path (score network -> DDPM reverse step / Langevin step).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it, and only as the checker or as the CPU baseline being
timed.  Nothing under `two-for-one-diffusion_b200/` imports it: the product
path has no CPU fallback and raises if the CUDA library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), but it is pure Python/PyTorch and imports in the build
container, so `oracle/make_golden.py` runs the UNMODIFIED reference from
/root/reference on fixed seeds and commits its inputs/outputs under
`tests/golden/`.  `tests/test_oracle_golden.py` checks every oracle function
against those fixtures (CPU, no GPU needed).

Modules
  score_ref.py      literal restatement of models/graph_transformer.py (autograd forces)
  collapsed_ref.py  the algebraically collapsed form + hand-written reverse mode that
                    the CUDA kernel implements (fp64-capable; validates the kernel math)
  sampler_ref.py    models/ddpm.py sampling half, dynamics/langevin*.py integrators, utils.py primitives
  weights.py        seeded synthetic state-dicts in the reference's checkpoint key layout
"""
