"""deepsent-b200: sm_100a kernels + host schedule for the Deep Sentiment training step of anthonyhu/tumblr-emotions."""
__version__ = "0.1.0"
