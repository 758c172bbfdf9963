"""Mirror of the reference's ``argParser.ArgsParser`` (argParser.py:3-72): same flags and defaults, plus
``--K``, ``--precision`` and ``--world_size`` for the batched B200 path."""
import argparse


class ArgsParser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--input_size', type=int, default=2)
    parser.add_argument('--output_size', type=int, default=5)
    parser.add_argument('--rnn_size', type=int, default=128, help='size of RNN hidden state')
    parser.add_argument('--num_layers', type=int, default=2)
    parser.add_argument('--model', type=str, default='lstm')
    parser.add_argument('--batch_size', type=int, default=16)
    parser.add_argument('--seq_length', type=int, default=12)
    parser.add_argument('--pred_len', type=int, default=12)
    parser.add_argument('--obs_len', type=int, default=8)
    parser.add_argument('--num_epochs', type=int, default=10)
    parser.add_argument('--save_every', type=int, default=50)
    parser.add_argument('--grad_clip', type=float, default=10.)
    parser.add_argument('--learning_rate', type=float, default=0.005)
    parser.add_argument('--decay_rate', type=float, default=0.95)
    parser.add_argument('--dropout', type=float, default=0.8)
    parser.add_argument('--embedding_size', type=int, default=64)
    parser.add_argument('--neighborhood_size', type=int, default=64)
    parser.add_argument('--grid_size', type=int, default=4)
    parser.add_argument('--num_freq_blocks', type=int, default=10)
    parser.add_argument('--maxNumPeds', type=int, default=20)
    parser.add_argument('--leaveDataset', type=int, default=2)
    parser.add_argument('--lambda_param', type=float, default=0.0005)
    # additions of the batched B200 path
    parser.add_argument('--K', type=int, default=20, help='samples per agent for best-of-K')
    parser.add_argument('--precision', type=str, default='fp16', choices=['fp32', 'fp16', 'bf16', 'bf16x3'])
    parser.add_argument('--world_size', type=int, default=1)
    parser.add_argument('--variant', type=str, default='mcr', choices=['mc', 'mcr'],
                        help='model of the batched path: g2k_lstm_mcr (the reference driver\'s, relational) or g2k_lstm_mc')
    parser.add_argument('--data_root', type=str, default=None, help='directory holding eth/ and ucy/')
    parser.add_argument('--save_dir', type=str, default=None,
                        help='directory for TensorFlow-bundle checkpoints (train.py:330-343); resumed from if it holds one')
    parser.add_argument('--max_agents', type=int, default=64, help='padded agents per scene (N)')
