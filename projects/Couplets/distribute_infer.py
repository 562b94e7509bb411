"""Multi-GPU variant of ``infer.py`` (reference projects/Couplets/distribute_infer.py): launch with
``bash tools/infer.sh projects/Couplets/distribute_infer.py <gpus>``; the parallel layout comes from the config."""
from projects.Couplets.infer import GeneratorForEager  # noqa: F401

if __name__ == "__main__":
    import sys

    from libai_b200.utils import distributed as dist

    gen = GeneratorForEager("projects/Couplets/configs/config.py", "output/couplet/model_final", "data_test/couplets/vocabs")
    out = gen.infer(sys.argv[1] if len(sys.argv) > 1 else "天增岁月人增寿")
    if dist.is_main_process():
        print(out)
