mkdir -p gpurun_out
for L in 1 2 3 6; do
B200SA_LIB=$PWD/build_exp/lib_look$L.so python tools/phase_times.py --kinds=dna,bytes 100000000 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('LOOK$L', d['input'], {k:v for k,v in d['phases_ms'].items() if 'lms' in k or k=='classify'})"
done
