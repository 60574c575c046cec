"""Recipe: stage the UNMODIFIED reference under ``oracle/_ref/`` so that the CPU baseline of record can run on
the GPU box (where /root/reference does not exist).

TEST / BENCH INFRASTRUCTURE.  ``oracle/_ref/`` is git-ignored (reference sources never enter the history) but NOT
gpurun-ignored, so the staged files travel with the snapshot like the built ``.so``.  The reference is pure Python
(no build step): "building" it is a byte-for-byte copy of the files the path needs, with a manifest of their
sha256 so a reader can check that nothing was edited.

    python -m oracle.build_ref            # copies when /root/reference is present; no-op otherwise

Run by ``__graft_entry__.build()`` in the development container.  Users: ``oracle/ref_shims.py`` (import shims),
``oracle/ref_baseline.py`` (bench.py --impl reference / cpu_baseline), ``oracle/gen_golden.py`` (fixtures).
"""
import hashlib
import json
import os
import shutil

SRC = os.environ.get("IC3NET_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# the files of the rollout / training path; StarCraft, plotting and rand.py stay behind
FILES = ["action_utils.py", "comm.py", "data.py", "env_wrappers.py", "main.py", "models.py", "multi_processing.py",
         "trainer.py", "utils.py", "LICENSE",
         "ic3net-envs/ic3net_envs/__init__.py", "ic3net-envs/ic3net_envs/predator_prey_env.py",
         "ic3net-envs/ic3net_envs/traffic_helper.py", "ic3net-envs/ic3net_envs/traffic_junction_env.py",
         "ic3net-envs/LICENSE.md"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose=False):
    """Copy the reference files into oracle/_ref (idempotent).  Returns the destination or None when the
    reference tree is not present (GPU box: the staged copy, if any, is used as is)."""
    if not os.path.isdir(os.path.join(SRC, "ic3net-envs", "ic3net_envs")):
        return DST if os.path.isdir(os.path.join(DST, "ic3net-envs", "ic3net_envs")) else None
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or _sha(s) != _sha(d):
            shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=SRC, files=manifest), f, indent=1, sort_keys=True)
    if verbose:
        print("staged %d reference files under %s" % (len(manifest), DST))
    return DST


if __name__ == "__main__":
    print(build(verbose=True))
