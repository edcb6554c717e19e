from voicemap_b200.utils import *  # noqa: F401,F403
from voicemap_b200.utils import (BatchPreProcessor, NShotEvaluationCallback, contrastive_loss,  # noqa: F401
                                 n_shot_task_evaluation, n_shot_task_evaluation_batched, preprocess_instances,
                                 whiten)
