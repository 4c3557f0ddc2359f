"""CPU oracle for the PCA-BO hot path (TEST INFRASTRUCTURE -- not the product).

This package restates, on the CPU, the algorithm that IvanBanny/para-ortho-pca-bo
executes on its hot path (rank weights -> weighted PCA -> GP caches -> log-EI ->
optimize_acqf -> inverse map), following the reference source line by line
(citations in every function are relative to /root/reference) and the published
semantics of its un-vendored third-party dependencies (botorch 0.13.0,
gpytorch 1.14, pyDOE 0.3.8, ioh 0.3.18 -- none installed here; scikit-learn,
scipy and torch ARE installed and are called directly where the reference
calls them).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package. The product (the C-ABI library under
para-ortho-pca-bo_b200/ and the Algorithms mirror) never does.

Pinning status
  * DoE (oracle.doe)                  : pinned bit-exactly on the reference's shipped .dat files
  * BBOB f15 / f20 (oracle.bbob)      : pinned on the reference's shipped .dat / .json files
  * BBOB f16 / f17 / f21-f24          : parity unpinned (no vectors in the reference)
  * weights / PCA / GP / logEI / acqf : parity unpinned by the reference (it has no tests and
                                        botorch/gpytorch cannot be installed here); sklearn PCA,
                                        scipy L-BFGS-B and torch SobolEngine are the real
                                        third-party code, the GP/logEI algebra is restated.
"""
