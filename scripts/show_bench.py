import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        r = d.get('roofline') or {}
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'lig/s', round(d['value'], 1), 'gcl_us', round(r.get('us_per_launch', 0), 1),
              {k: round(v, 3) for k, v in (r.get('step_share_ms') or {}).items()})
    except Exception as e:
        print(f, 'unreadable:', e)
