#!/bin/bash
# ThreadSanitizer and Address/UB-Sanitizer runs of the host layer's structural-edit path over the recording double (no device needed).
#   bash tests/abi_recorder/sanitize.sh [bodies]  (needs the EnTT / GLM headers: PHYSECS_ENTT_INCLUDE / PHYSECS_GLM_INCLUDE or /root/reference)
set -e
cd "$(dirname "$0")/../.."
REF=${PHYSECS_REFERENCE:-/root/reference}
ENTT=${PHYSECS_ENTT_INCLUDE:-$REF/vendor/entt-3.12.2/single_include/entt}
GLM=${PHYSECS_GLM_INCLUDE:-"$REF/vendor/glm 0.9.9.8"}
OUT=tests/abi_recorder/_build
mkdir -p $OUT
for SAN in thread address,undefined; do
    BIN=$OUT/sanitize_${SAN%%,*}
    g++ -std=c++17 -O1 -g -fsanitize=$SAN -fno-omit-frame-pointer -ffp-contract=off -DGLM_FORCE_INLINE -I include/Physecs -I include/Physecs/Joints -I include \
        -I "$GLM" -I "$ENTT" physecs_b200/host/{Scene,Meshes,MassUtil,scene_harness}.cpp tests/abi_recorder/pb_recorder.cpp physecs_b200/csrc/trimesh_build.cpp tests/abi_recorder/sanitize_driver.cpp -o $BIN -lpthread
    echo "== -fsanitize=$SAN"
    TSAN_OPTIONS=halt_on_error=1 ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 $BIN ${1:-30000}
    BIN=$OUT/sanitize_batch_${SAN%%,*}
    g++ -std=c++17 -O1 -g -fsanitize=$SAN -fno-omit-frame-pointer -I tests/abi_recorder/stub physecs_b200/csrc/batch.cpp tests/abi_recorder/pb_recorder.cpp physecs_b200/csrc/trimesh_build.cpp \
        tests/abi_recorder/sanitize_batch_driver.cpp -o $BIN -lpthread
    TSAN_OPTIONS=halt_on_error=1 ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 $BIN
done
