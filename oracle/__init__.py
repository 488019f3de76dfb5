"""CPU oracle for the MSDA hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(HankerSia/Apollo-Vision-Net, ``projects/mmdet3d_plugin/bevformer/modules``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product package
``apollo-vision-net_b200`` never does and fails loudly without its CUDA library.

Parity pinning: the reference ships no golden vectors for this path
(SURVEY.md section 8c).  The oracle is pinned instead against outputs of the
reference's *own unmodified module code* executed in the build container
through ``oracle/refshim`` (an mmcv stand-in that lets the reference files
import); the resulting vectors are committed under ``tests/golden/`` together
with ``tests/golden/make_golden.py`` that produced them.
"""
