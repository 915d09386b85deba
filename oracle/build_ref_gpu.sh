#!/usr/bin/env bash
# Builds oracle/_ref/pt_ref_gpu: the reference's own Renderer.cpp + Scene.cpp compiled as CUDA for sm_100a (its kernels as written),
# the GPU-side baseline of bench.py (`reference_gpu`).  TEST / MEASUREMENT INFRASTRUCTURE, never part of the product.
# Sources are read from /root/reference, copied to a throw-away directory outside the repository and patched there:
#   P1  utility.h:44        drop the duplicate `inline` (does not compile otherwise)
#   P3  Renderer.cpp:359    `return false;` on the path that falls off the end (what the device code does anyway)
#   K   Config.h            GRID_*/RESOLUTION_*/ITER -> __managed__ ints defined there (one translation unit; read by host and device),
#                           depth literal (Renderer.cpp:550) -> __managed__ int
#   C   Renderer.cpp:604,617  ray counter next to the two closest-hit launches
#   S   Renderer.cpp:567-648  `err = cudaDeviceSynchronize();` -> skipped when ptap_cfg_sync == 0 (second timing variant)
set -euo pipefail
REF=${PTAP_REFERENCE:-/root/reference}/PathTracerAP
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -f "$REF/Renderer.cpp" ]; then echo "reference sources not found under $REF" >&2; exit 3; fi
TMP=$(mktemp -d /tmp/ptap_refgpu_build.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
cp "$REF"/*.h "$REF"/*.cpp "$TMP"/
sed -i '44s/^inline unsigned int/unsigned int/' "$TMP/utility.h"                                   # P1
sed -i '604s/$/ ptap_rays_traced += nrays;/; 617s/$/ ptap_rays_traced += nrays;/' "$TMP/Renderer.cpp"   # C (before P3: line numbers of the reference)
sed -i '567,648s/err = cudaDeviceSynchronize();/err = ptap_cfg_sync ? cudaDeviceSynchronize() : cudaSuccess;/' "$TMP/Renderer.cpp"   # S
sed -i '359s/^    }$/    }\n    return false;/' "$TMP/Renderer.cpp"                               # P3
sed -i -E 's/^#define GRID_X .*$/__managed__ int ptap_cfg_grid_x = 25;\n#define GRID_X ptap_cfg_grid_x/;
           s/^#define GRID_Y .*$/__managed__ int ptap_cfg_grid_y = 25;\n#define GRID_Y ptap_cfg_grid_y/;
           s/^#define GRID_Z .*$/__managed__ int ptap_cfg_grid_z = 25;\n#define GRID_Z ptap_cfg_grid_z/;
           s/^#define RESOLUTION_X .*$/__managed__ int ptap_cfg_res_x = 1000;\n#define RESOLUTION_X ptap_cfg_res_x/;
           s/^#define RESOLUTION_Y .*$/__managed__ int ptap_cfg_res_y = 800;\n#define RESOLUTION_Y ptap_cfg_res_y/;
           s/^#define ITER .*$/__managed__ int ptap_cfg_iter = 500;\n#define ITER ptap_cfg_iter/' "$TMP/Config.h"
printf '\n__managed__ int ptap_cfg_depth = 5;\nextern long long ptap_rays_traced;\nextern int ptap_cfg_sync;\n' >> "$TMP/Config.h"
sed -i 's/meta_data\.remaining_bounces = 5;/meta_data.remaining_bounces = ptap_cfg_depth;/' "$TMP/Renderer.cpp"
sed -i 's/^private:/public:/' "$TMP/Scene.h"
grep -q 'ptap_rays_traced += nrays' "$TMP/Renderer.cpp"
[ "$(grep -c 'ptap_rays_traced += nrays' "$TMP/Renderer.cpp")" = 2 ]
grep -q 'computeRaySceneIntersectionKernel.*ptap_rays_traced' "$TMP/Renderer.cpp"
grep -q 'ptap_cfg_sync ? cudaDeviceSynchronize' "$TMP/Renderer.cpp"
grep -q 'return false;$' <(sed -n '360p' "$TMP/Renderer.cpp")
mkdir -p "$OUT"
nvcc -x cu -std=c++17 -O3 -w -gencode arch=compute_100a,code=sm_100a -lineinfo \
    -I"$HERE/../integration/shim" -I"$TMP" -I"$REF/external/include" \
    "$HERE/ref_gpu_main.cu" -o "$OUT/pt_ref_gpu"
echo "built $OUT/pt_ref_gpu"
