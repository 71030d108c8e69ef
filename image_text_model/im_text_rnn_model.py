"""Drop-in module path of the reference's image_text_model/im_text_rnn_model.py: the same names resolve to the B200 engine
(`tumblr_emotions_b200.api`).  Callers such as parallel_computing/job_train.py:23-31 and job_correlation.py:4-6 keep working:

    from image_text_model.im_text_rnn_model import train_deep_sentiment, correlation_matrix
"""
from tumblr_emotions_b200.api import (DEEP_SENTIMENT_CONFIG as _CONFIG, DeepSentiment, _POST_SIZE, correlation_matrix,  # noqa: F401
                                      train_deep_sentiment)
