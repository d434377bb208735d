#!/usr/bin/env python
"""Synthetic input files for the reference's main.cpp pipeline (SURVEY.md App. E formats), from one seed.

    python tools/make_main_inputs.py OUT_DIR [--users 900] [--seed 7]

Writes into OUT_DIR: cluster.conf, tweets.tsv (-d), coins_queries.csv, vader_lexicon.csv, proj2_input.csv.
main.cpp reads ./cluster.conf from its working directory (main.cpp:49), so run the binary with cwd = OUT_DIR.
"""
import argparse
import os

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--users", type=int, default=900)
    ap.add_argument("--coins", type=int, default=100)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--P", type=int, default=20)
    ap.add_argument("--clusters", type=int, default=12)
    ap.add_argument("--proj2-clusters", type=int, default=40)
    ap.add_argument("--iterations", type=int, default=3)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    os.makedirs(a.out, exist_ok=True)
    D = a.coins
    # coins: 5 name variants per line, the 5th (index 4) is the printed one (main.cpp:563-566); a few short lines
    with open(os.path.join(a.out, "coins_queries.csv"), "w") as f:
        for c in range(D):
            names = ["coin%d" % c, "c%d" % c, "C%d" % c, "$c%d" % c, "Coin-%d" % c]
            if c % 17 == 3:
                names = names[:2]
            f.write("\t".join(names) + "\n")
    words = ["w%d" % i for i in range(400)]
    with open(os.path.join(a.out, "vader_lexicon.csv"), "w") as f:
        for w in words:
            f.write("%s\t%.1f\n" % (w, rng.uniform(-3.5, 3.5)))
    # tweets: user_id \t tweet_id \t words...   (first line: label \t P)
    lines = ["tweets\t%d" % a.P]
    tweet_ids = []
    tid = 0
    taste = rng.integers(0, D, size=(a.users, 6))
    for u in range(a.users):
        for _ in range(int(rng.integers(1, 7))):
            toks = []
            for _ in range(int(rng.integers(2, 9))):
                toks.append(words[int(rng.integers(0, len(words)))] if rng.random() < 0.85 else "xx%d" % int(rng.integers(0, 50)))
            for _ in range(int(rng.integers(0, 4))):
                c = int(taste[u, int(rng.integers(0, 6))])
                variants = 2 if c % 17 == 3 else 5
                v = int(rng.integers(0, variants))
                toks.append(["coin%d", "c%d", "C%d", "$c%d", "Coin-%d"][v] % c)
            order = rng.permutation(len(toks))
            lines.append("\t".join(["%d" % (1000 + u), "%d" % tid] + [toks[i] for i in order]))
            tweet_ids.append(tid)
            tid += 1
    with open(os.path.join(a.out, "tweets.tsv"), "w") as f:
        f.write("\n".join(lines) + "\n")
    # project-2 tweet vectors: id, then non-negative "tf-idf"-like coordinates around proj2 cluster centres
    d2 = 203   # wider than 128: the tweet vectors of the original data set are of this kind
    # sparse centres (12 active coordinates each): cosine distances between a tweet and its cluster mean stay well
    # above min_dist_kmeans, so the first k_means call replaces the centres (main.cpp:109-110 frees them afterwards)
    centres = np.zeros((a.proj2_clusters, d2))
    for c in range(a.proj2_clusters):
        centres[c, rng.choice(d2, 12, replace=False)] = rng.gamma(4.0, 1.0, size=12)
    with open(os.path.join(a.out, "proj2_input.csv"), "w") as f:
        for t in tweet_ids:
            if t % 23 == 5:
                continue  # a tweet without a vector
            v = centres[int(rng.integers(0, a.proj2_clusters))] * rng.uniform(0.6, 1.4, size=d2) + rng.gamma(1.0, 0.5, size=d2) * (rng.random(d2) < 0.08)
            f.write("%d,%s\n" % (t, ",".join("%.6f" % x for x in v)))
    with open(os.path.join(a.out, "cluster.conf"), "w") as f:
        f.write("proj_2_input ./proj2_input.csv\nproj_2_csv_delimiter ,\nproj_2_number_of_clusters %d\n" % a.proj2_clusters)
        f.write("number_of_clusters %d\nnumber_of_hash_functions 4\nnumber_of_hash_tables 5\ncsv_delimiter 9\n" % a.clusters)
        f.write("lsh_bucket_div 100\neuclidean_h_w 0.4\nmax_algo_iterations %d\nmin_dist_kmeans 0.05\n" % a.iterations)
        f.write("lexicon_file ./vader_lexicon.csv\nquery_file ./coins_queries.csv\n")
    print("wrote %d tweets of %d users, %d coins to %s" % (tid, a.users, D, a.out))


if __name__ == "__main__":
    main()
