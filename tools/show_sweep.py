import glob, json
for f in sorted(glob.glob('gpurun_out/sweep_*.json')):
    try:
        d = json.load(open(f))
        ri = d['resident']
        print(f.split('sweep_')[1][:-5], round(d['value'] / 1e6, 2), 'M/s', round(d['ms_per_step'], 1), 'ms frac', round(d['roofline']['frac'], 4), ri['last_S'], ri['last_I'], ri['last_J'], ri['hbm'])
    except Exception as e:
        print(f, 'ERR', e)
