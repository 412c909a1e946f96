"""Random-gather ceiling of the device: independent vs dependent aligned 32/64/128 B loads
over a table far larger than L2.  Writes one JSON line per configuration."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import genedex_b200 as gdx

lib = gdx._lib.load()
out = []
for table_gb in (1.5, 12.0):
    for rec in (32, 64, 128):
        for chained in (0, 1):
            g, l = C.c_double(), C.c_double()
            rc = lib.gdx_measure_random_gather(0, int(table_gb * (1 << 30)), rec, 1 << 31 if not chained else 1 << 30,
                                               chained, C.byref(g), C.byref(l))
            assert rc == 0, lib.gdx_last_error_message()
            line = dict(table_gb=table_gb, record_bytes=rec, chained=bool(chained), gbps=round(g.value, 1),
                        gloads_per_s=round(l.value, 2))
            print(json.dumps(line))
            out.append(line)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
