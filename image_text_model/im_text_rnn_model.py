"""Drop-in module path of the reference's image_text_model/im_text_rnn_model.py: the same names resolve to the B200 engine
(`tumblr_emotions_b200.api`).  Callers such as parallel_computing/job_train.py:23-31, job_correlation.py:4-6 and
job_evaluate.py:20-31 keep working:

    from image_text_model.im_text_rnn_model import train_deep_sentiment, correlation_matrix, evaluate_deep_sentiment
"""
from tumblr_emotions_b200.api import (DEEP_SENTIMENT_CONFIG as _CONFIG, DeepSentiment, _POST_SIZE, correlation_matrix,  # noqa: F401
                                      day_of_week_trend, evaluate_deep_sentiment, outliers_detection, train_deep_sentiment,
                                      word_most_relevant)
from tumblr_emotions_b200.text_preprocessing import _load_embedding_weights_glove, _paragraph_to_ids  # noqa: F401
