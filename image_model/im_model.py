"""Drop-in module path of the reference's image_model/im_model.py (ImageModel :139-164, train_image_model :166-225,
evaluate_image_model :227-262, get_init_fn :118-137), backed by `tumblr_emotions_b200.api`."""
from tumblr_emotions_b200.api import (IMAGE_CONFIG as _CONFIG, ImageModel, _RANDOM_SEED, evaluate_image_model, get_init_fn,  # noqa: F401
                                      train_image_model)
