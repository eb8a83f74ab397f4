#!/bin/bash
# Round evidence with the kept tap tables as the default: full GPU suite, smoke, bench line, reference arm, launch list,
# ncu --set full of the sweep kernel.
bash profiles/runners/gpu_round2.sh r4b
