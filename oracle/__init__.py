"""CPU oracle for the anytime voxel-decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.

PARITY UNPINNED: the reference (bogus2000/anytime-3D-reconstruction) ships no
tests, golden vectors, seeds or saved weights for this path, and its arithmetic
lives in an un-vendored, unpinned third-party dependency (TensorFlow 2.x
``tf.keras``; not installable here).  This oracle therefore restates the
published Keras semantics of the layers the reference calls, following the
reference call sites line by line:

* decoder graph        src/net_core/autoencoder3D.py:41-70,104-139
* sampling             src/module/function.py:35-38
* voxelPrecisionRecall src/module/function.py:100-115
* mask / fills         src/module/nolbo.py:1472-1510 (and :431-439)
* K-sample mean        src/module/nolbo_test.py:167-177

and is cross-checked internally: an fp64 numpy restatement written directly
from the *definition* of a Keras ``Conv3DTranspose`` (the adjoint of a
SAME-padded strided forward convolution) pins the torch-fp32 restatement, and
the committed fixtures in ``tests/golden`` pin both against regressions.
"""
