run() { echo "== $*"; env "$@" timeout -s KILL 120 python bench.py --steps 3 --warmup 2 --seconds 2 --skip-cpu --e2e-recordings 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,1), 'G ch-samp/s  k1_ms', round(d['roofline']['kernel_ms'],1))"; }
run OFP_K1_WS=0
run OFP_K1_WS=1 OFP_K1_STAGES=2
run OFP_K1_WS=1 OFP_K1_STAGES=3
run OFP_K1_WS=1 OFP_K1_STAGES=4
run OFP_K1_WS=1 OFP_K1_STAGES=2 OFP_K1_WS_TILECAP=44
run OFP_K1_WS=1 OFP_K1_STAGES=3 OFP_K1_NDB=4 OFP_K1_SLACK=2

