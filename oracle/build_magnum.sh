#!/bin/bash
# Builds the reference's own vendored Corrade + GL-less Magnum + CgltfImporter/StbImageImporter (contrib/, unmodified,
# out-of-tree: /root/reference is read-only) into $PREFIX — the link dependencies of the reference's
# src/mesh_tools/{consolidate,compute_tangents}.cpp, which oracle/build_ref.py compiles into oracle/_ref/meshtool.
# Probe-verified recipe of SURVEY.md Appendix D. Test infrastructure only.
set -e
REF=${SLB_REFERENCE:-/root/reference}
WORK=${SLB_REF_WORK:-/tmp/slb_ref_build}
PREFIX=$WORK/install
COMMON="-GNinja -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release -DCMAKE_INSTALL_PREFIX=$PREFIX -DCMAKE_PREFIX_PATH=$PREFIX -DBUILD_STATIC=ON -DBUILD_STATIC_PIC=ON -DBUILD_PLUGINS_STATIC=ON"
mkdir -p $WORK
cmake -S $REF/contrib/corrade -B $WORK/corrade $COMMON -DWITH_INTERCONNECT=OFF -DWITH_TESTSUITE=OFF > $WORK/corrade.log 2>&1
ninja -C $WORK/corrade install >> $WORK/corrade.log 2>&1
cmake -S $REF/contrib/magnum -B $WORK/magnum $COMMON -DWITH_GL=OFF -DTARGET_GL=OFF -DWITH_SHADERS=OFF -DWITH_DEBUGTOOLS=OFF \
      -DWITH_TEXT=OFF -DWITH_TEXTURETOOLS=OFF -DWITH_SHADERTOOLS=OFF -DWITH_AUDIO=OFF -DWITH_MESHTOOLS=ON -DWITH_PRIMITIVES=ON \
      -DWITH_SCENEGRAPH=ON -DWITH_TRADE=ON -DWITH_ANYIMAGEIMPORTER=ON > $WORK/magnum.log 2>&1
ninja -C $WORK/magnum install >> $WORK/magnum.log 2>&1
cmake -S $REF/contrib/magnum-plugins -B $WORK/plugins $COMMON -DWITH_CGLTFIMPORTER=ON -DWITH_STBIMAGEIMPORTER=ON \
      -DWITH_STANFORDIMPORTER=ON > $WORK/plugins.log 2>&1
ninja -C $WORK/plugins install >> $WORK/plugins.log 2>&1
echo "$PREFIX"
