"""Installs the UNMODIFIED reference modules into git-ignored baseline/_ref/ (they travel to the GPU box with the
snapshot): the reference has no installable package (its pyproject.toml lists build requirements only), so the
"install" is a verbatim copy of /root/reference/src/*.py plus a manifest of their sha256 digests.  Used by
  * bench.py --impl reference   (the reference's own Fingerprint.quantize under multiprocessing.Pool), and
  * tests/test_reference_dropin_gpu.py (the reference's Database / search_db running on the faiss shim).
Nothing under dctdomain_b200/ imports it.  No-op when /root/reference is absent (GPU box: the copy already exists)."""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('DCTD_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')


def main():
    src = os.path.join(SRC, 'src')
    if not os.path.isdir(src):
        print(f'{src} not present: baseline/_ref left as it is')
        return 0
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(src)):
        if not name.endswith('.py'):
            continue
        shutil.copyfile(os.path.join(src, name), os.path.join(DST, name))
        manifest[name] = hashlib.sha256(open(os.path.join(DST, name), 'rb').read()).hexdigest()
    json.dump({'source': src, 'files': manifest}, open(os.path.join(DST, 'MANIFEST.json'), 'w'), indent=1)
    print(f'installed {len(manifest)} reference modules into {DST}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
