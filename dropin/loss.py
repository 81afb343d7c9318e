"""shim: lets the reference train.py import the B200 implementation under the reference module name"""
from dcvgan_b200.loss import *  # noqa: F401,F403
