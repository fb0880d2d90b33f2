#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== NACC default" > gpurun_out/exp.log
python tools/trace_conv.py --only "128 " >> gpurun_out/exp.log 2>&1
python tools/trace_conv.py --only "D 512->512 333 @9x288" >> gpurun_out/exp.log 2>&1
echo "== NACC=2" >> gpurun_out/exp.log
CVVAE_CONV_NACC=2 python tools/trace_conv.py --only "128 " >> gpurun_out/exp.log 2>&1
echo "== NACC=1 (N=256)" >> gpurun_out/exp.log
CVVAE_CONV_NACC=1 python tools/trace_conv.py --only "D 512->512 333 @9x288" >> gpurun_out/exp.log 2>&1
CVVAE_CONV_NACC=1 CVVAE_CONV_CTA_GROUP=1 python tools/trace_conv.py --only "D 512->512 333 @9x288" >> gpurun_out/exp.log 2>&1
