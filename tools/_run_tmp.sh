VARIANTS="L256 D128 L256 D128" CONFIGS="2 4" bash tools/sweep_variants.sh
RAPPAS_B200_LIB=build/variants/D128.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
