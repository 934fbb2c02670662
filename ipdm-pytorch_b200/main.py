"""Command-line entry point of the B200 build.  Same contract as the reference's script (options parsed by
Config.default_config, the progressive_domain_denoiser evaluation loop):

    python main.py --load_option_path Config/Mayo-Config/test_progressive_option.json --convertor FBP [--test_batch_size 16]

Only the test modes exist on this path; training modes fail loudly inside the denoiser's constructor.
"""
import sys


def run(argv=None):
    import Config.default_config as config
    import Utils.train_test_utils as runner
    options = config.default_cfg(argv)
    runner.progressive_domain_denoiser(options).fit()
    return 0


if __name__ == '__main__':
    sys.exit(run(sys.argv[1:]))
