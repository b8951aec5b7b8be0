"""CPU oracle for the deep-image-matching hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement (torch fp32 / numpy) of the reference's
per-pair hot path: SuperPoint / ALIKED extraction, LightGlue / LighterGlue / SuperGlue
matching and the kornia brute-force descriptor matcher.  It exists only to check the CUDA product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* the product package (``deep-image-matching_b200`` / ``dim_b200``) never
  imports it and fails loudly when its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
* SuperPoint  - pinned: checked against the reference's vendored model code
  (thirdparty/SuperGluePretrainedNetwork/models/superpoint.py) executed in the
  authoring container; vectors in tests/golden/ (oracle/gen_golden.py).
* LightGlue   - pinned the same way against thirdparty/LightGlue/lightglue/
  lightglue.py with deterministic seeded weights (no pretrained LightGlue
  checkpoint exists offline) and CUDA control-flow semantics.
* ALIKED      - pinned: bit-identical to the reference's LightGlue port of ALIKED (thirdparty/LightGlue/lightglue/aliked.py)
  with the vendored aliked-n16rot checkpoint on crops of the reference's test photo.
* LightGlue with TRAINED weights - pinned: the vendored LighterGlue checkpoint (LightGlue architecture, 96 / 1 head / 6 layers) on
  XFeat features of the reference's test photos; expected outputs from the reference's LightGlue class.
* SuperGlue   - pinned: against thirdparty/SuperGluePretrainedNetwork/models/superglue.py with the vendored trained outdoor
  checkpoint (checked at generation time) and seeded weights (stored vectors).
* kornia NN   - PARITY UNPINNED: kornia 0.8.1 is not vendored under the
  reference nor installable here; match_nn/mnn/snn/smnn restate its published
  algorithm.  A secondary check uses hloc's mutual-NN
  (thirdparty/hloc/matchers/nearest_neighbor.py) for the cosine variant.
"""
