#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_keras_backend.py -q -p no:cacheprovider 2>&1 | tail -8
timeout 400 python tools/lazy_adam_stream.py > gpurun_out/lazy_adam_stream.jsonl 2> gpurun_out/lazy_adam_stream.err; echo "stream rc=$?"; cat gpurun_out/lazy_adam_stream.jsonl | cut -c1-700; tail -3 gpurun_out/lazy_adam_stream.err
timeout 500 python tools/train_e2e.py 16 > gpurun_out/train_e2e_ring.json 2> gpurun_out/train_e2e_ring.err; echo "e2e rc=$?"; cat gpurun_out/train_e2e_ring.json; tail -3 gpurun_out/train_e2e_ring.err
C2V_BATCH_RING=0 timeout 500 python tools/train_e2e.py 16 > gpurun_out/train_e2e_noring.json 2> gpurun_out/train_e2e_noring.err; echo "e2e noring rc=$?"; cat gpurun_out/train_e2e_noring.json
