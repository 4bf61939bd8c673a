#!/bin/bash
tools/profile.sh 200000 7 final_2e5
tools/profile.sh 1000000 7 final_1e6
tools/gpu_final.sh
