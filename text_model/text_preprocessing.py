"""Drop-in module path of the reference's text_model/text_preprocessing.py for the two functions the models import
(image_text_model/im_text_rnn_model.py:18): the GloVe loader :13-35 and _paragraph_to_ids :84-105."""
from tumblr_emotions_b200.text_preprocessing import _PUNCTUATION, _load_embedding_weights_glove, _paragraph_to_ids  # noqa: F401
