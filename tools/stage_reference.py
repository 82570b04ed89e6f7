"""Stage the UNMODIFIED reference files that the timed / end-to-end comparisons need under baseline/_ref/ (git-ignored, but
shipped to the GPU box with the gpurun snapshot; /root/reference itself does not exist there).

    python tools/stage_reference.py [--reference /root/reference]

Nothing is edited: the files are byte-for-byte copies (the script prints their sha256), and nothing of the product
imports them -- bench.py's reference legs, tests/test_reference_e2e.py and tools/run_configs.py do."""
import argparse
import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [
    "concept_vit/similarity.py",                 # the scoring functions (the path this repo replaces)
    "concept_vit/CLIP_og_utils.py",              # hook factory + get_similarity_from_activations (describe_clip_neurons)
    "concept_vit/utils.py",
    "concept_vit/og_utils.py",
    "concept_vit/describe_clip_neurons.py",      # the driver that is run unchanged
    "model/modules/efficientnet_custom.py",      # the reference's EfficientNet class (config c2: random-init B5)
    "model/modules/efficient_net_custom_utils.py",
    "Concepts/Specific_concepts_sorted.txt",     # the 763-concept set
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    dst_root = os.path.join(ROOT, "baseline", "_ref")
    for rel in FILES:
        src = os.path.join(args.reference, rel)
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            print("%s  %s" % (hashlib.sha256(f.read()).hexdigest()[:16], rel))
    print("staged under", dst_root)


if __name__ == "__main__":
    main()
