"""Command-line switch surface of the reference drivers (options.py:8-116 of the reference), kept
flag-for-flag so launch scripts (slurm_files/multimodal/*.sh) carry over.  Differences: the parser is
table-driven, takes an optional argv, and accepts the hyphenated spellings `--features-only` /
`--cross-attention-only` that the reference README and SLURM scripts use but its argparse rejects
(SURVEY.md §0)."""
import argparse

_B = argparse.BooleanOptionalAction

# (flags, kwargs) — defaults identical to the reference
_SPEC = [
    (("--epochs",), dict(type=int, default=100, help="number of rounds of training")),
    (("--dataset_folder_name",), dict(type=str, default="", help="dataset folder name in the base location")),
    (("--dataset_folder_name_val",), dict(type=str, default="", help="val dataset folder name")),
    (("--lr",), dict(type=float, default=0.001, help="learning rate")),
    (("--image_text_dropout",), dict(type=float, default=0.33, help="chance of dropping either text or image")),
    (("--image_prob_dropout",), dict(type=float, default=0.7, help="chance the dropped modality is the image")),
    (("--reg",), dict(type=float, default=1e-2, help="regularization rate")),
    (("--model_dropout",), dict(type=float, default=0.6, help="model FC layer dropout")),
    (("--tl",), dict(action=_B, default=True, help="use transfer learning")),
    (("--balance_weights",), dict(action=_B, default=False, help="use class balance weights")),
    (("--ft_epochs",), dict(type=int, default=15, help="number of fine tuning epochs")),
    (("--fraction_lr",), dict(type=float, default=5, help="LR divisor for fine tuning")),
    (("--image_model",), dict(type=str, default="b4", help="model name")),
    (("--text_model",), dict(type=str, default="distilbert", help="model name")),
    (("--model_path",), dict(type=str, default="", help="checkpoint to evaluate")),
    (("--acc_steps",), dict(type=int, default=0, help="gradient accumulation steps")),
    (("--acc_steps_FT",), dict(type=int, default=0, help="gradient accumulation steps (fine tuning)")),
    (("--num_neurons_FC",), dict(type=int, default=256, help="neurons in FC layers")),
    (("--batch_size",), dict(type=int, default=16, help="batch size")),
    (("--batch_size_FT",), dict(type=int, default=16, help="batch size for fine tuning")),
    (("--opt",), dict(type=str, default="sgd", help="optimizer")),
    (("--base_path",), dict(type=str, default="", help="base path for saved models")),
    (("--calculate_dataset_stats",), dict(action=_B, default=False, help="compute normalization stats")),
    (("--prob_aug",), dict(type=float, default=0.6, help="probability of applying augmentations")),
    (("--late_fusion",), dict(type=str, default="gated", help="late fusion strategy (MM_RCA | hierarchical | ...)")),
    (("--label_smoothing",), dict(type=float, default=0.0, help="label smoothing fraction")),
    (("--name",), dict(type=str, help="run description")),
    (("--reverse",), dict(action=_B, default=False, help="use RCA or not")),
    (("--features_only", "--features-only"), dict(action=_B, default=False, dest="features_only",
                                                  help="use only the extracted features")),
    (("--cross_attention_only", "--cross-attention-only"), dict(action=_B, default=False,
                                                                dest="cross_attention_only",
                                                                help="use only the cross attention features")),
    (("--extended_desc_train",), dict(type=str, help="extended description train CSV")),
    (("--extended_desc_val",), dict(type=str, help="extended description val CSV")),
    (("--balanced_sampler",), dict(action=_B, default=False, help="use balanced sampler")),
    (("--use_synonyms",), dict(action=_B, default=False, help="synonymizer augmentation for text")),
    (("--prob_aug_text",), dict(type=float, default=0.6, help="prob of text synonymization")),
    (("--classifier_weights",), dict(type=str, help="classifier head weights of the Q-Former model")),
    # B200 extensions (not in the reference)
    (("--compute",), dict(type=str, default="fp32", choices=("fp32", "bf16"),
                          help="head arithmetic: fp32 SIMT (1e-4 contract) or bf16 tensor cores (2e-2 contract)")),
]


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser()
    for flags, kw in _SPEC:
        parser.add_argument(*flags, **kw)
    return parser


def args_parser(argv=None):
    return build_parser().parse_args(argv)
