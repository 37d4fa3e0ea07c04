#!/bin/bash
# every h16_probe test in its own process (an illegal instruction poisons the CUDA context)
for t in split_fp16 split_bf16 split_fp16_noaug split_bf16_noaug split_fp16_swap orbh drain32 drain16; do
  timeout 30 easysfm_b200/bin/h16_probe $t 2>&1 | tail -6
done
