for v in 0 1 2 3; do echo -n "TVS_LN_FWD=$v "; TVS_LN_FWD=$v timeout 60 tests/native/selftest lnprof 50 2>&1 | grep -E "layernorm fwd|FAILED" | head -2; done
