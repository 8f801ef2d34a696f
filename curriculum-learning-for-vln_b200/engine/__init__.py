"""Training engine (reference: src/engine/)."""
from .optim import FlatOptimizer, build_optimizer
from .trainer import ClassicTrainer, NaiveCurriculum, SelfPacedCurriculum, TrainStep, build_trainer
from .evaluator import Evaluation, evaluate
from .graphs import GraphedTrainStep

__all__ = ["FlatOptimizer", "build_optimizer", "ClassicTrainer", "NaiveCurriculum", "SelfPacedCurriculum", "TrainStep",
           "build_trainer", "Evaluation", "evaluate", "GraphedTrainStep"]
