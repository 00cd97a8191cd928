# duration of the FIRST k_build_ws launch (identical inputs whatever the variant computes) for each library given
for lib in "$@"; do
  echo -n "$lib: "
  VISFS_BA_LIB=$lib timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:${KREGEX:-k_build_ws} -c 1 python tools/profile_c3.py 512 2>/dev/null | grep -E "gpu__time_duration" | awk '{print $(NF-1), $NF}'
done
