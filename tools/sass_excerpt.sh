#!/bin/bash
# Regenerate profiles/r02_conv_engine.sass: the tcgen05 / TMA / TMEM instructions of the tensor-core kernels (needs no GPU).
cd "$(dirname "$0")/.." && cuobjdump -sass packnet_sfm_b200/libpacknet_b200.so | c++filt | grep -E "Function :|UTCHMMA|UTMALDG|UBLKCP|LDTM|UTCBAR|UTCATOMSWS"
