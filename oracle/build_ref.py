"""Stage the UNMODIFIED reference hot path under oracle/_ref/ so that it travels to the GPU box.  TEST / BASELINE
INFRASTRUCTURE ONLY (oracle/_ref/ is git-ignored build output; nothing under scene_generation_b200/ may import it).

The reference is a pure-Python program: "building" it means placing the package files it needs for
the path `train.py:190-215` — scene_generation/{args,bilinear,discriminators,generators,graph,layers,layout,losses,
metrics,model,trainer,utils}.py and data/{__init__,utils}.py, byte for byte — where `oracle/ref_harness.py` can import them when /root/reference
does not exist (the GPU box).  bench.py's reference arm (`--impl reference`: the reference's own train loop on the
host cores; `--impl reference-gpu`: the same loop through stock PyTorch eager / cuDNN on the B200) runs from here.
A MANIFEST with the sha256 of every staged file is written next to them; `verify()` re-checks it.

    python -m oracle.build_ref            # in the build container (needs /root/reference)
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('SG_REFERENCE_SRC', '/root/reference')
DST = os.path.join(HERE, '_ref')
FILES = ['__init__.py', 'args.py', 'bilinear.py', 'discriminators.py', 'generators.py', 'graph.py', 'layers.py',
         'layout.py', 'losses.py', 'metrics.py', 'model.py', 'trainer.py', 'utils.py',
         'data/__init__.py', 'data/utils.py']        # trainer.py:8 imports imagenet_deprocess_batch (logging only)


def _sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def build():
    """copy the files (unmodified) and write the manifest; returns DST, or None when the reference tree is absent"""
    src_pkg = os.path.join(SRC, 'scene_generation')
    if not os.path.isdir(src_pkg):
        return None
    dst_pkg = os.path.join(DST, 'scene_generation')
    os.makedirs(dst_pkg, exist_ok=True)
    manifest = {}
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(dst_pkg, f)), exist_ok=True)
        shutil.copyfile(os.path.join(src_pkg, f), os.path.join(dst_pkg, f))
        manifest[f] = _sha(os.path.join(dst_pkg, f))
    json.dump({'source': src_pkg, 'sha256': manifest}, open(os.path.join(DST, 'MANIFEST.json'), 'w'), indent=1)
    return DST


def staged():
    return os.path.isfile(os.path.join(DST, 'MANIFEST.json'))


def verify():
    """the staged files still have the recorded hashes (i.e. are the reference's bytes)"""
    m = json.load(open(os.path.join(DST, 'MANIFEST.json')))['sha256']
    return all(_sha(os.path.join(DST, 'scene_generation', f)) == h for f, h in m.items())


if __name__ == '__main__':
    print(build() or 'reference tree not found at %s' % SRC)
