#!/bin/bash
bash tools/gpu_round2_j.sh
bash tools/gpu_round2_k.sh
