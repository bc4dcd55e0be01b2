// Dataset.h -- drop-in for MetaGenomics/Dataset.h:20-51 on top of libogb's host Dataset stage
// (ogb_dataset_*): same constructor and accessors, same read IDs (rank in the lexicographic order of
// the canonical strands + 1, Dataset.cpp:316-345).
#ifndef OGB_HOST_DATASET_H_
#define OGB_HOST_DATASET_H_

#include "Common.h"
#include "Read.h"

class Dataset
{
	private:
		UINT64 numberOfReads;
		UINT64 numberOfUniqueReads;
		UINT64 minimumOverlapLength;
		vector<Read *> *reads;
		ogb_dataset *store;								// packed, sorted, unique reads (host)
		ogb_context *mateContext;						// device context for the batched mate-pair pass (NULL: host loop)
		bool storeMatePairInformation(string fileName, UINT64 minOverlap, UINT64 datasetNumber);
		void adopt(UINT64 minOverlap);					// finalize `store` and create the Read objects

	public:
		vector<string> pairedEndDatasetFileNames;
		vector<string> singleEndDatasetFileNames;
		UINT64 shortestReadLength;
		UINT64 longestReadLength;

		Dataset(void);
		Dataset(vector<string> pairedEndFileNames, vector<string> singleEndFileNames, UINT64 minOverlap);
		// In-memory variant (not in the reference): n reads, read i = bases[offsets[i]..offsets[i+1]).
		Dataset(const char *bases, const uint64_t *offsets, UINT64 n, UINT64 minOverlap);
		~Dataset(void);
		UINT64 getNumberOfReads(void);
		UINT64 getNumberOfUniqueReads(void);
		bool printDataset(void);
		Read *getReadFromString(const string &read);
		Read *getReadFromID(UINT64 ID);
		void readMatePairsFromFile(void);
		// Not in the reference: the mate-pair pass (storeMatePairInformation) runs as one batched lookup on this context's GPU,
		// which must hold this data set's reads, index and containment marks (OverlapGraph sets it before OverlapGraph.cpp:142).
		void setMatePairContext(ogb_context *ctx) { mateContext = ctx; }
		void saveReads(string fileName);

		string readString(UINT64 ID, int strand) const;	// used by Read
		const ogb_dataset *handle() const { return store; }
};

#endif
