#!/bin/bash
# forward staging-ring variant: build it first with
#   GG_VARIANT_UNITS=blend.cu python tools/build_variant.py ring -DGG_FWD_RING=1
# result: profiles/r02_blend_bwd_variants.txt (measured and killed)
mkdir -p gpurun_out
export GG_LIB_PATH=$PWD/gaussiangrasper_b200/variants/libgg_ring.so
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "fused or golden or config1 or small or blend or raster" > gpurun_out/t_ring.log 2>&1; echo "ring tests rc=$?"; tail -2 gpurun_out/t_ring.log
unset GG_LIB_PATH
bash tools/gpu_variants_stage.sh
