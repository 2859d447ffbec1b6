#!/bin/bash
for c in 8 24 32 64 100000; do echo "HM_UMMA_CHUNK=$c"; HM_UMMA_CHUNK=$c ENC_AB_CFG=B HM_ENC_MODE=3 timeout 200 python tools/enc_ab.py child; done
