"""CPU oracle for the CGVAE equivariant message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``coarsegrainingvae_b200/`` imports this
package.  The only allowed users are ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there only as the
checker / the CPU arm -- never as the thing measured or shipped as the product.

The oracle is a functional (state-dict driven) restatement, in plain PyTorch CPU
ops and numpy, of the reference algorithm in
``/root/reference/CoarseGrainingVAE/{modules,conv,cgvae,data}.py`` and the loss
assembly in ``/root/reference/scripts/utils.py``.  Every function cites the
reference file:line it follows.

Parity pinning: the reference ships no tests / golden vectors,
so the oracle is pinned against outputs of the reference itself, imported
read-only in the build container by ``oracle/ref_shim.py`` and frozen into
``tests/golden/*.npz`` by ``tests/golden/make_golden.py`` (committed).  The
``-m "not gpu"`` tests check the oracle against those fixtures (forward values and
autograd gradients), and -- when ``/root/reference`` is present -- against the live
reference as well.
"""
