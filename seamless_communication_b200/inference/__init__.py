from .generator import SequenceGeneratorOptions, UnitYGenerator
from .translator import BatchedSpeechOutput, Modality, Task, Translator
