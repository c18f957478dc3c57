for rep in 1 2; do
for t in base walk minb4 minb2; do
if [ $t = base ]; then unset B200SA_LIB; else export B200SA_LIB=$PWD/build_exp/lib_$t.so; fi
python tools/phase_times.py --kinds=dna 100000000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); p=d['phases_ms']; print('$t', p['classify'], p['lms_sort'], p['lms_groups'])"
done; done
