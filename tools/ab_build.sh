# A/B of libvisfs_ba variants on the resident C3 batch: bash tools/ab_build.sh lib1.so lib2.so ...
for lib in "$@"; do
  for rep in 1 2; do
    echo -n "$lib: "; VISFS_BA_LIB=$lib timeout 200 python tools/profile_c3.py 512 | tail -1
  done
done
