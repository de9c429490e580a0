"""Compare the SASS of every kernel in one .cu file between a git revision and the working tree (no GPU needed).

    python scripts/sass_diff.py <git-rev> tpgsr_b200/csrc/conv_flat.cu

Used to prove that an edit which adds template instantiations / experimental switches leaves the instruction streams of the GPU-validated
default kernels byte-for-byte unchanged (kernels are matched by their demangled name with defaulted template arguments stripped)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo']


def sass(obj):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['cu++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', cur)
            cur = re.sub(r'\(int\)', '', cur)
            cur = re.sub(r'<(\d+), 1>', r'<\1>', cur)                  # a defaulted trailing template argument = the old instantiation
            res[cur] = []
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*;', line)
        if m and cur:
            res[cur].append(m.group(1))
    return res


def main():
    rev, path = sys.argv[1], sys.argv[2]
    src_dir = os.path.join(ROOT, os.path.dirname(path))
    with tempfile.TemporaryDirectory() as d:
        old_cu = os.path.join(d, 'old_' + os.path.basename(path))
        with open(old_cu, 'w') as f:
            f.write(subprocess.run(['git', 'show', f'{rev}:{path}'], cwd=ROOT, capture_output=True, text=True, check=True).stdout)
        subprocess.run(NVCC + ['-I', src_dir, '-c', old_cu, '-o', os.path.join(d, 'old.o')], check=True)
        subprocess.run(NVCC + ['-I', src_dir, '-c', os.path.join(ROOT, path), '-o', os.path.join(d, 'new.o')], check=True)
        old, new = sass(os.path.join(d, 'old.o')), sass(os.path.join(d, 'new.o'))
    bad = 0
    for name, code in sorted(new.items()):
        short = name.split('(')[0].replace('void ', '')[:70]
        if name not in old:
            print(f'NEW        {len(code):5d}  {short}')
        elif old[name] == code:
            print(f'IDENTICAL  {len(code):5d}  {short}')
        else:
            bad += 1
            print(f'DIFFERENT  {len(code):5d}  {short}   (was {len(old[name])})')
    for name in sorted(set(old) - set(new)):
        print(f'REMOVED           {name[:70]}')
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
