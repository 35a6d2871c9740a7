from .vocoder import Vocoder, load_vocoder_model
