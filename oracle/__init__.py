"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's variational-layer hot path
(yliess86/BayeFormers, mounted read-only at /root/reference while this repo is
being built).  Nothing in `bayeformers_b200/` (the product) imports this
package.  The only permitted importers are:

  * `tests/`                      -- as the parity checker,
  * `__graft_entry__.smoke()`     -- as the checker of the one smoke invocation,
  * `bench.py`                    -- the `cpu_baseline` leg and `--impl reference`.

Parity status: **pinned by execution of the reference**.  The reference ships
no tests, golden vectors or fixtures of its own, so the
oracle is pinned against outputs of the unmodified reference imported from
/root/reference in the build container; the generating script is
`tests/golden/make_golden.py` and the fixtures it wrote are committed under
`tests/golden/`.  `tests/test_oracle_golden.py` re-checks the oracle against
those fixtures on every run (CPU only), and -- when /root/reference is present
-- against the live reference as well.
"""
