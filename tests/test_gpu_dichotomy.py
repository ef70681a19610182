"""The reference's only engine-level acceptance test, DichotomyTest.Main (ProbQA/PqaCoreTests/DichotomyTest.cpp:10-101),
re-expressed over the batch entry points: 1000 questions x 5 answers x 1000 targets, initAmount 0.1; a quiz asks
NextQuestion, answers by the +-32 band rule, RecordAnswer, ListTopTargets(10), up to 100 questions, stops when the hidden
target is listed, then RecordQuizTarget + ReleaseQuiz. After 3 000 000 questions asked, 10 000 further trials must list
the target in the top 10 at least 98 % of the time. The reference runs quizzes one after another; here 1024 quizzes run
concurrently (like the reference's multi-threaded PqaClient.cpp:238-245), which changes the interleaving of training,
not the statistics."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def band_answers(questions, guesses):
    q, g = np.asarray(questions), np.asarray(guesses)
    return np.where(g < q - 32, 0, np.where(g < q, 1, np.where(g == q, 2, np.where(g <= q + 32, 3, 4)))).astype(np.int64)


def test_dichotomy_main():
    from probqa_b200 import engine as pqa
    Q, K, T, B = 1000, 5, 1000, 1024
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), rng_seed=20171016,
                                                    initial_quiz_capacity=B)
    rng = np.random.default_rng(1)
    n_trials_wanted, max_len, top = 10 * 1000, 100, 10
    n_trials = n_correct = 0
    quizzes = eng.start_quiz_batch(B)
    guesses = rng.integers(0, T, size=B)
    lengths = np.zeros(B, dtype=np.int64)
    counted = np.zeros(B, dtype=bool)          # quiz started after the 3M mark => counts as a trial
    while n_trials < n_trials_wanted:
        questions = eng.next_question_batch(quizzes)
        assert np.all(questions >= 0)
        eng.record_answer_batch(quizzes, band_answers(questions, guesses))
        lengths += 1
        items, counts = eng.list_top_targets_batch(quizzes, top)
        assert np.all(counts == top)
        hit = np.any(items["iTarget"] == guesses[:, None], axis=1)
        done = hit | (lengths >= max_len)
        if not done.any():
            continue
        idx = np.nonzero(done)[0]
        n_correct += int(np.sum(hit[idx] & counted[idx]))
        n_trials += int(np.sum(counted[idx]))
        eng.record_quiz_target_batch(quizzes[idx], guesses[idx])
        eng.release_quiz_batch(quizzes[idx])
        fresh = eng.start_quiz_batch(idx.size)
        quizzes[idx] = fresh
        guesses[idx] = rng.integers(0, T, size=idx.size)
        lengths[idx] = 0
        counted[idx] = eng.get_total_questions_asked() > 3 * 1000 * 1000
    assert n_correct >= 0.98 * n_trials, (n_correct, n_trials)
    print("dichotomy: %d / %d trials listed the target in the top %d (%.2f %%), %d questions asked" % (
        n_correct, n_trials, top, 100.0 * n_correct / n_trials, eng.get_total_questions_asked()))
