"""One NJ+TopHits step through vft_nj_build (for ncu launch lists / captures). argv: taxa [columns]"""
import sys
sys.path.insert(0, '.')
from veryfasttree_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
chars = synth.make_alignment(n, L, 'nt', 1)
chars = chars[synth.unique_rows(chars)]
t = api.nj_build(api.encode(chars, 'nt'), 4, 32, trace=False)
print('taxa', chars.shape[0], 'launches', t.stats['counters']['launches'], 'device ms', t.stats['deviceMsResident'])
