#!/bin/bash
# gy access pattern (B=16, R=384 channels): L=3000 (conv1) and 1500 (conv2)
P=./tools/probe/stream_probe
for L in 3000 1500; do
  $P 16 384 $L 64 3 1 3 2       # current gy kernel: stage = 192 rows x 32 windows, 3 stages, 2 CTAs/SM
  $P 16 384 $L 64 3 1 4 2       # 4 stages
  $P 16 384 $L 64 3 1 2 4       # 2 stages, 4 CTAs/SM
  $P 16 384 $L 64 3 1 4 3       # 4 stages, 3 CTAs... (72 KB x 3)
  $P 16 384 $L 64 6 1 2 2       # stage = whole 384-row tile, 2 stages
  $P 16 384 $L 32 3 2 3 2       # stage = 96 rows x 64 windows
  $P 16 384 $L 16 3 4 3 2       # stage = 48 rows x 128 windows
  $P 16 384 $L 64 1 3 3 2       # stage = 64 rows x 96 windows
  $P 16 384 $L 192 1 1 3 2      # one 192-row box per stage
  $P 16 384 $L 64 3 1 3 2 0     # no L2 promotion
  $P 16 384 $L 64 3 1 3 2 3     # 256B promotion -> (3 = L2_256B, 2 = L2_128B)
  $P 16 384 $L 64 1 1 8 2       # 8 KB stages x 8
  $P 16 384 $L 64 1 1 12 2      # 8 KB stages x 12
done
# x of conv2 as the forward reads it is different (72-wide unswizzled boxes); here: 128-window wide tiles, 32-row boxes
$P 16 384 3000 32 1 4 3 4
$P 16 384 3000 32 2 4 3 2
