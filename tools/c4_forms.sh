#!/bin/bash
# C4 (64 frames, 32768x16384 ComplexF32): A-form pair of GEMMs against the Gram-form GEMM, with the phase times of the batched driver
RLS_TRACE_BATCH=1 python tools/run_configs.py c4 > gpurun_out/c4_forms.jsonl 2> gpurun_out/c4_forms.err
