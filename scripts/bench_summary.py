"""One-line digest of a bench.py JSON line (GPU-call scripts print it so the numbers survive in the call's stdout)."""
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r, rc, eg, inf, par = d.get('roofline', {}), d.get('roofline_crnn', {}), d.get('gpu_eager_baseline', {}), d.get('inference', {}), d.get('parity', {})
    print(sys.argv[2], 'cfg', d['config']['config_id'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']),
          'launches', d['launches_per_step'], '| conv us', r.get('launch_us'), 'frac', r.get('frac'), 'flushed', r.get('launch_us_flushed'), r.get('frac_flushed'), r.get('error'), '| crnn conv us', rc.get('launch_us'), 'frac', rc.get('frac'), rc.get('error'),
          '| eager', eg.get('value'), eg.get('inference_images_per_s'), eg.get('error'), '| infer', inf.get('value'), inf.get('error'),
          '| parity', par.get('sr_rel_l2_vs_oracle'), par.get('psnr'), par.get('ssim'), par.get('error'),
          '| cpu', d.get('cpu_baseline', {}).get('value'), '| loss', d['config']['last_loss'], '| clocks', d.get('clocks'))
except Exception as e:
    print(sys.argv[2], 'no line:', e)
