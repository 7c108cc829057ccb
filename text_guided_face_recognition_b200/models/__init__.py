"""Drop-in mirror of the reference's `models` package for the FCAM hot path.

Put this package's parent directory first on sys.path and the reference drivers'
`from models.losses import sent_loss, words_loss, ...` / `from models import metrics, losses`
(src/train_encoders_bert.py:19,25) resolve here.
"""
