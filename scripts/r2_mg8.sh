#!/bin/bash
cd "$(dirname "$0")/.."
bash scripts/r2_mg.sh 8 quick
bash scripts/r2_mg.sh 4 quick
