"""Entry point with the reference's main.py contract (main.py:1-6):
    python main.py --load_option_path Config/Mayo-Config/test_progressive_option.json --convertor FBP
"""
from Config.default_config import default_cfg
from Utils.train_test_utils import progressive_domain_denoiser

if __name__ == '__main__':
    opt = default_cfg()
    model = progressive_domain_denoiser(opt)
    model.fit()
