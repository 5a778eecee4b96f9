for cfg in "B200LC_CUHD_VARIANT=1 B200LC_CUHD_MULTI_BITS=12" "B200LC_CUHD_VARIANT=2" "B200LC_CUHD_VARIANT=2 B200LC_CUHD_MULTI_BITS=12" "B200LC_CUHD_VARIANT=1 B200LC_CUHD_MULTI_BITS=11"; do
    echo "=== $cfg"
    env $cfg timeout 60 python tools/bench_paths.py cuhd --mib 1024 | grep -o '"decode_ms": [0-9.]*'
    env $cfg timeout 60 python tools/bench_paths.py cuhd --mib 64 | grep -o '"decode_ms": [0-9.]*'
done
