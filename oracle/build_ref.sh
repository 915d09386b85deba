#!/usr/bin/env bash
# Builds oracle/_ref/libptap_ref.so: the reference's own Renderer.cpp + Scene.cpp compiled
# for the host CPU (recipe: SURVEY.md Appendix A).  TEST INFRASTRUCTURE.
#
# The reference sources are read where they lie (/root/reference, read-only), copied into a
# throw-away directory OUTSIDE the repository, patched there, and only the shared library is
# written to oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).  The reference's
# own build system (a Windows .vcxproj) is not used.
#
# Patches applied to the temporary copy (none changes arithmetic):
#   P1  utility.h:44       drop the duplicate `inline` specifier (does not compile otherwise)
#   P2  Renderer.cpp:577.. `k<<<blocks,threads>>>(...)` -> LAUNCH(k, blocks, threads, ...)
#   P3  Renderer.cpp:359   `return false;` on the bbox-miss path that falls off the end (UB;
#                          nvcc device code behaves as false - SURVEY.md 0.5)
#   K   Config.h           GRID_*/RESOLUTION_*/ITER macros -> runtime ints; depth literal
#                          (Renderer.cpp:550) -> runtime int; Scene.h private -> public
#   I   Renderer.cpp:210,395  probe statements recording (model, triangle, t, u, v)
set -euo pipefail
REF=${PTAP_REFERENCE:-/root/reference}/PathTracerAP
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -f "$REF/Renderer.cpp" ]; then echo "reference sources not found under $REF" >&2; exit 3; fi
TMP=$(mktemp -d /tmp/ptap_ref_build.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
cp "$REF"/*.h "$REF"/*.cpp "$TMP"/

# P1
sed -i '44s/^inline unsigned int/unsigned int/' "$TMP/utility.h"
# P2
sed -i -E 's/([A-Za-z_]+) *<< *<(blocks), *(threads) *>> *> *\(/LAUNCH(\1, \2, \3, /' "$TMP/Renderer.cpp" "$TMP/Experimentation.h"
# I (before P3 so that line numbers still match the reference)
sed -i '395s/$/ if (ptap_probe_out) ptap_probe_out[iray] = PtapProbe{imodel, ptap_probe_tri, ptap_probe_t, ptap_probe_u, ptap_probe_v};/' "$TMP/Renderer.cpp"
sed -i '210s/$/ ptap_probe_tri = itriangle; ptap_probe_t = t; ptap_probe_u = u; ptap_probe_v = v;/' "$TMP/Renderer.cpp"
# P3
sed -i '359s/^    }$/    }\n    return false;/' "$TMP/Renderer.cpp"
# K
sed -i -E 's/^#define GRID_X .*$/extern int ptap_cfg_grid_x;\n#define GRID_X ptap_cfg_grid_x/;
           s/^#define GRID_Y .*$/extern int ptap_cfg_grid_y;\n#define GRID_Y ptap_cfg_grid_y/;
           s/^#define GRID_Z .*$/extern int ptap_cfg_grid_z;\n#define GRID_Z ptap_cfg_grid_z/;
           s/^#define RESOLUTION_X .*$/extern int ptap_cfg_res_x;\n#define RESOLUTION_X ptap_cfg_res_x/;
           s/^#define RESOLUTION_Y .*$/extern int ptap_cfg_res_y;\n#define RESOLUTION_Y ptap_cfg_res_y/;
           s/^#define ITER .*$/extern int ptap_cfg_iter;\n#define ITER ptap_cfg_iter/' "$TMP/Config.h"
printf '\nextern int ptap_cfg_depth;\n' >> "$TMP/Config.h"
sed -i 's/meta_data\.remaining_bounces = 5;/meta_data.remaining_bounces = ptap_cfg_depth;/' "$TMP/Renderer.cpp"
sed -i 's/^private:/public:/' "$TMP/Scene.h"
# probe declarations visible to Renderer.cpp
cat >> "$TMP/Config.h" <<'EOT'
struct PtapProbe;
extern thread_local int ptap_probe_tri;
extern thread_local float ptap_probe_t, ptap_probe_u, ptap_probe_v;
extern PtapProbe* ptap_probe_out;
EOT

# sanity: each patch must have landed
grep -q 'LAUNCH(computeRaySceneIntersectionKernel' "$TMP/Renderer.cpp"
grep -q 'return false;$' <(sed -n '360p' "$TMP/Renderer.cpp")
grep -q 'ptap_probe_tri = itriangle' "$TMP/Renderer.cpp"
grep -q 'ptap_probe_out\[iray\]' "$TMP/Renderer.cpp"
grep -q 'ptap_cfg_depth;' "$TMP/Renderer.cpp"
! grep -q '<< *<' "$TMP/Renderer.cpp"

mkdir -p "$OUT"
# -ffp-contract=off is REQUIRED: it is what makes host results independent of -O level,
# -march and thread count; the product kernels compute the hit-deciding
# arithmetic un-contracted for the same reason.
g++ -std=c++17 -O2 -ffp-contract=off -fopenmp -fPIC -shared -w \
    -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_OMP \
    -I"$HERE/shim" -I"$HERE/../integration/shim" -I"$TMP" -I"$REF/external/include" -I/usr/local/cuda/include \
    -x c++ "$HERE/ref_harness.cpp" -o "$OUT/libptap_ref.so"
echo "built $OUT/libptap_ref.so"
