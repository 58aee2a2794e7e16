#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (facebookresearch/hanabi_SAD)
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored, but it
# travels to the GPU box with the gpurun snapshot).  Recipe = SURVEY.md section 8(c): the
# reference's own CMake does not configure in this image (ABI flag vs torch 2.11, FindPythonInterp,
# pybind11 2.3 vs Python 3.12), so the handful of translation units are compiled directly.
#
# Outputs (all under oracle/_ref/):
#   rela<EXT>.so, hanalearn<EXT>.so   the reference pybind modules (asserts kept: no -DNDEBUG)
#   libhanabi.a                       the reference HLE simulator (no torch)
#   pyhanabi/                         the reference's python side (r2d2.py gets the one-token
#                                     TorchScript fix `s.dim()` -> `priv_s.dim()`, r2d2.py:69)
# Nothing here is imported by the product; only tests/, smoke() and bench.py's reference /
# cpu_baseline legs use it.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF not present (GPU box?) -- using prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/pyhanabi"
PY=${PYTHON:-python}
TORCH=$($PY -c "import torch,os;print(os.path.dirname(torch.__file__))")
EXT=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
HLE="$REF/hanabi-learning-environment/hanabi_lib"
cd "$OUT/obj"
# 1. HLE (no torch): the 9 files listed in hanabi_lib/CMakeLists.txt:1
for f in hanabi_card hanabi_game hanabi_hand hanabi_history_item hanabi_move hanabi_observation hanabi_state util canonical_encoders; do
  if [ ! -f $f.o ] || [ "$HLE/$f.cc" -nt $f.o ]; then
    g++ -O3 -std=c++14 -fPIC -w -c "$HLE/$f.cc" -o $f.o &
  fi
done
wait
rm -f ../libhanabi.a
ar rcs ../libhanabi.a hanabi_card.o hanabi_game.o hanabi_hand.o hanabi_history_item.o hanabi_move.o hanabi_observation.o hanabi_state.o util.o canonical_encoders.o
# 2. torch TUs (C++17 required by torch 2.11; torch's bundled pybind11 comes from $TORCH/include)
FL="-O3 -std=c++17 -fPIC -w -I$REF -I$TORCH/include -I$TORCH/include/torch/csrc/api/include -I$PYINC"
[ -f rela_pybind.o ] || g++ $FL -c "$REF/rela/pybind.cc" -o rela_pybind.o &
[ -f transition.o ]  || g++ $FL -c "$REF/rela/transition.cc" -o transition.o &
[ -f hanabi_env.o ]  || g++ $FL -c "$REF/cpp/hanabi_env.cc" -o hanabi_env.o &
[ -f hl_pybind.o ]   || g++ $FL -c "$REF/cpp/pybind.cc" -o hl_pybind.o &
wait
LIBS="-L$TORCH/lib -ltorch -ltorch_cpu -lc10 -ltorch_python -Wl,-rpath,$TORCH/lib"
g++ -shared -o "../rela$EXT" rela_pybind.o transition.o $LIBS
g++ -shared -o "../hanalearn$EXT" hl_pybind.o hanabi_env.o ../libhanabi.a $LIBS
# 3. python side (generated copy, never committed)
cp -r "$REF/pyhanabi/common_utils" "$OUT/pyhanabi/" 2>/dev/null || true
for f in utils.py create.py eval.py set_path.py selfplay.py; do cp "$REF/pyhanabi/$f" "$OUT/pyhanabi/$f"; done
sed 's/% s\.dim()/% priv_s.dim()/' "$REF/pyhanabi/r2d2.py" > "$OUT/pyhanabi/r2d2.py"
mkdir -p "$OUT/pyhanabi/tools"
for f in eval_model.py convert_model.py obl_model.py action_matrix.py; do cp "$REF/pyhanabi/tools/$f" "$OUT/pyhanabi/tools/$f"; done
echo "[build_ref] done -> $OUT"
