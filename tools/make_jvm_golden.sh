#!/bin/bash
# Produces tests/golden_jvm/*.json with the REAL RAPPAS (the pin of SURVEY.md 8c) -- to be run on any machine
# that has a JDK >= 8, the reference checkout and fastutil-8.2.2.jar (absent from the reference tree,
# .MISSING_LARGE_BLOBS:1).  Neither this repository's build image nor its GPU box has a JVM.
#   RAPPAS_SRC=/path/to/RAPPAS FASTUTIL_JAR=/path/to/fastutil-8.2.2.jar tools/make_jvm_golden.sh
set -euo pipefail
: "${RAPPAS_SRC:?set RAPPAS_SRC to the reference checkout}"; : "${FASTUTIL_JAR:?set FASTUTIL_JAR}"
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=$HERE/tests/golden_jvm; mkdir -p "$OUT" "$HERE/build/jvm"
CP="$FASTUTIL_JAR:$RAPPAS_SRC/lib/json_simple-1.1.jar:$RAPPAS_SRC/lib/Jacksum.jar"
# 1. the inputs: the same seeded DBs / reads the GPU parity tests use, as .rgdb + FASTA  (python, no GPU needed)
python "$HERE/tests/golden_jvm/make_inputs.py" "$OUT"
# 2. the reference + the two harness classes
find "$RAPPAS_SRC/src" -name '*.java' > "$HERE/build/jvm/sources.txt"
javac -nowarn -d "$HERE/build/jvm" -cp "$CP" @"$HERE/build/jvm/sources.txt" \
      "$HERE/integration/java/tools/RgdbImporter.java" "$HERE/integration/java/tools/GoldenDump.java"
# 3. one run per case
for db in "$OUT"/*.rgdb; do
  base=${db%.rgdb}
  java -Xmx8g -cp "$HERE/build/jvm:$CP" tools.GoldenDump "$db" "$base.fasta" "$base.json"
  java -Xmx8g -cp "$HERE/build/jvm:$CP" tools.GoldenDump "$db" "$base.fasta" "$base.ambmax.json" 7 0.01 true
done
echo "commit tests/golden_jvm/*.json (+ .rgdb / .fasta) -- tests/test_golden_jvm.py picks them up"
