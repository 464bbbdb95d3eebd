#!/bin/bash
# GPU call: correctness of attention2 on odd shapes, the attention unit tests, then the A/B timing against generation 1 and flash-attn 2
mkdir -p gpurun_out
timeout 300 python tools/attn_bench.py small > gpurun_out/attn_small.log 2>&1; echo "small rc=$?"
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k attention > gpurun_out/attn_pytest.log 2>&1; echo "pytest-attn rc=$?"
timeout 600 python tools/attn_bench.py vit llm mha gemm > gpurun_out/attn_big.log 2>&1; echo "big rc=$?"
(ls /usr/lib/x86_64-linux-gnu/ | grep -i -E "nvcuvid|nvidia-encode|nvjpeg"; python -c "import ctypes; print(ctypes.CDLL('libnvcuvid.so.1'))"; ffmpeg -version | head -1; python -c "import cv2; print(cv2.__version__)") > gpurun_out/nvdec_probe.log 2>&1
tail -5 gpurun_out/attn_small.log gpurun_out/attn_pytest.log; cat gpurun_out/attn_big.log | tail -30; cat gpurun_out/nvdec_probe.log
