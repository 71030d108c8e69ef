"""Drop-in module path of the reference's text_model/text_embedding.py (TextModel :37-86, train_text_model :89-150,
evaluate_text_model :152-187), backed by `tumblr_emotions_b200.api`."""
from tumblr_emotions_b200.api import TEXT_CONFIG as _CONFIG, TextModel, _POST_SIZE, evaluate_text_model, train_text_model  # noqa: F401
