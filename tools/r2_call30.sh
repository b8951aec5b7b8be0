#!/bin/bash
# GPU call 30: conv2a on CTA pairs as the default: SuperPoint users at odd sizes (feature store, tile preselection, low-res pairs, plugin).
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_store_and_sets.py tests/test_gpu_parity.py -m gpu -q -k "store or tile_preselection or image_set or lowres or superpoint_golden or superpoint_plugin or superpoint_other or nms_is_exact or pipeline" 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_pair2_default_tests.log
